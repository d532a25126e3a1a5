"""T5 relative positional encoding -- B200-native drop-in for the producer of the attention `bias`
(/root/reference/src/utils/positional_encoding.py:10-110, class RelativePositionalEncoding).

Same constructor, attributes (`relative_attention_bias` is the same `nn.Embedding(num_buckets, n_heads)`, so
checkpoints load unchanged), `_relative_position_bucket`, `compute_bias(query_length, key_length, device=None)`
and `forward(q, k, v) -> (q, k, v, bias)`.

What changes underneath: the reference materialises an int64 (M, N) bucket matrix, gathers an fp32 (M, N, H)
tensor, permutes it and casts it; here the bucket of every relative distance is computed once (with the
reference's own formula, so the buckets are bit-identical) into a small lookup table, and one CUDA kernel
writes the dense (1, H, M, N) bias directly in the requested dtype.  The backward is one segmented-sum kernel
(the reference's embedding backward scatter-add).  ALiBi / RoPE / FIRE are out of scope.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import _cabi

__all__ = ["RelativePositionalEncoding", "t5_bias_fwd", "t5_bias_bwd"]


@torch.library.custom_op("b200t5::t5_bias_fwd", mutates_args=(), device_types="cuda")
def t5_bias_fwd(table: torch.Tensor, lut: torch.Tensor, lut_zero: int, ctx_pos: Optional[torch.Tensor],
                mem_pos: Optional[torch.Tensor], M: int, N: int, out_dtype: torch.dtype) -> torch.Tensor:
    """table (num_buckets, H), lut int32 (bucket of rel + lut_zero) -> bias (1, H, M, N) in out_dtype."""
    _cabi.require_cuda(table, lut, ctx_pos, mem_pos)
    lib = _cabi.load()
    table = table.contiguous()
    nb, H = table.shape
    bias = torch.empty((1, H, M, N), dtype=out_dtype, device=table.device)
    rc = lib.b200t5_t5_bias_fwd(table.data_ptr(), lut.data_ptr(), lut_zero, lut.numel(),
                                ctx_pos.data_ptr() if ctx_pos is not None else None,
                                mem_pos.data_ptr() if mem_pos is not None else None, bias.data_ptr(), H, M, N, nb,
                                _cabi.dtype_code(table.dtype), _cabi.dtype_code(out_dtype), table.device.index,
                                _cabi.stream_ptr(table.device))
    _cabi.check(rc, "b200t5_t5_bias_fwd")
    return bias


@torch.library.register_fake("b200t5::t5_bias_fwd")
def _t5_bias_fwd_fake(table, lut, lut_zero, ctx_pos, mem_pos, M, N, out_dtype):
    return torch.empty((1, table.shape[1], M, N), dtype=out_dtype, device=table.device)


@torch.library.custom_op("b200t5::t5_bias_bwd", mutates_args=(), device_types="cuda")
def t5_bias_bwd(dbias: torch.Tensor, lut: torch.Tensor, lut_zero: int, ctx_pos: Optional[torch.Tensor],
                mem_pos: Optional[torch.Tensor], num_buckets: int) -> torch.Tensor:
    """dbias (1, H, M, N) -> dtable (num_buckets, H) fp32."""
    _cabi.require_cuda(dbias, lut, ctx_pos, mem_pos)
    lib = _cabi.load()
    dbias = dbias.contiguous()
    _, H, M, N = dbias.shape
    dtable = torch.empty((num_buckets, H), dtype=torch.float32, device=dbias.device)
    rc = lib.b200t5_t5_bias_bwd(dbias.data_ptr(), lut.data_ptr(), lut_zero, lut.numel(),
                                ctx_pos.data_ptr() if ctx_pos is not None else None,
                                mem_pos.data_ptr() if mem_pos is not None else None, dtable.data_ptr(), H, M, N,
                                num_buckets, _cabi.dtype_code(dbias.dtype), dbias.device.index,
                                _cabi.stream_ptr(dbias.device))
    _cabi.check(rc, "b200t5_t5_bias_bwd")
    return dtable


@torch.library.register_fake("b200t5::t5_bias_bwd")
def _t5_bias_bwd_fake(dbias, lut, lut_zero, ctx_pos, mem_pos, num_buckets):
    return torch.empty((num_buckets, dbias.shape[1]), dtype=torch.float32, device=dbias.device)


class _T5Bias(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, lut, lut_zero, ctx_pos, mem_pos, M, N, out_dtype):
        bias = torch.ops.b200t5.t5_bias_fwd(table, lut, lut_zero, ctx_pos, mem_pos, M, N, out_dtype)
        ctx.save_for_backward(lut, ctx_pos, mem_pos)
        ctx.lut_zero = lut_zero
        ctx.num_buckets = table.shape[0]
        ctx.table_dtype = table.dtype
        return bias

    @staticmethod
    def backward(ctx, dbias):
        lut, ctx_pos, mem_pos = ctx.saved_tensors
        dtable = torch.ops.b200t5.t5_bias_bwd(dbias, lut, ctx.lut_zero, ctx_pos, mem_pos, ctx.num_buckets)
        return dtable.to(ctx.table_dtype), None, None, None, None, None, None, None


class RelativePositionalEncoding(nn.Module):

    def __init__(self, relative_attention_num_buckets, relative_attention_max_distance, n_heads, max_sequence_length,
                 bidirectional=True, randomized_position=False):
        super().__init__()
        self.relative_attention_num_buckets = relative_attention_num_buckets
        self.relative_attention_max_distance = relative_attention_max_distance
        self.n_heads = n_heads
        self.max_sequence_length = max_sequence_length
        self.bidirectional = bidirectional
        self.randomized_position = randomized_position
        self.relative_attention_bias = nn.Embedding(self.relative_attention_num_buckets, self.n_heads)
        self._lut_cache = {}

    @staticmethod
    def _relative_position_bucket(relative_position, bidirectional=True, num_buckets=32, max_distance=128):
        """relative position (memory - query) -> bucket in [0, num_buckets); the reference formula (:25-71),
        same operations in the same order and dtypes so the integer results are identical."""
        relative_buckets = 0
        if bidirectional:
            num_buckets //= 2
            relative_buckets += (relative_position > 0).to(torch.long) * num_buckets
            relative_position = torch.abs(relative_position)
        else:
            relative_position = -torch.min(relative_position, torch.zeros_like(relative_position))
        max_exact = num_buckets // 2
        is_small = relative_position < max_exact
        relative_position_if_large = max_exact + (
            torch.log(relative_position.float() / max_exact)
            / torch.log(torch.tensor(max_distance / max_exact))
            * (num_buckets - max_exact)
        ).to(torch.long)
        relative_position_if_large = torch.min(
            relative_position_if_large, torch.full_like(relative_position_if_large, num_buckets - 1)
        )
        relative_buckets += torch.where(is_small, relative_position, relative_position_if_large)
        return relative_buckets

    def _bucket_lut(self, lo: int, hi: int, device):
        """int32 lookup table over the relative positions lo..hi (inclusive), evaluated on `device` like the
        reference evaluates its (M, N) bucket matrix there."""
        key = (lo, hi, str(device))
        lut = self._lut_cache.get(key)
        if lut is None:
            rel = torch.arange(lo, hi + 1, dtype=torch.long, device=device)
            lut = self._relative_position_bucket(rel, bidirectional=self.bidirectional,
                                                 num_buckets=self.relative_attention_num_buckets,
                                                 max_distance=self.relative_attention_max_distance).to(torch.int32)
            self._lut_cache[key] = lut
        return lut

    def compute_bias(self, query_length, key_length, device=None, dtype=None):
        """Binned relative position bias, (1, n_heads, query_length, key_length).  `dtype` (extension): emit the bias
        directly in that dtype instead of the table's dtype followed by a cast."""
        weight = self.relative_attention_bias.weight
        if device is None:
            device = weight.device
        out_dtype = dtype if dtype is not None else weight.dtype
        if self.randomized_position:
            # same sampling as the reference (:78-87): sorted random subsets of 0..max_sequence_length-1, rooted at 0
            context_indices_rand, _ = torch.sort(torch.randperm(self.max_sequence_length)[:query_length])
            context_indices_rand[0] = 0
            memory_indices_rand, _ = torch.sort(torch.randperm(self.max_sequence_length)[:key_length])
            memory_indices_rand[0] = 0
            ctx_pos = context_indices_rand.to(device=device, dtype=torch.int32)
            mem_pos = memory_indices_rand.to(device=device, dtype=torch.int32)
            lo, hi = -(self.max_sequence_length - 1), self.max_sequence_length - 1
        else:
            ctx_pos = mem_pos = None
            lo, hi = -(query_length - 1), key_length - 1
        lut = self._bucket_lut(lo, hi, device)
        return _T5Bias.apply(weight, lut, -lo, ctx_pos, mem_pos, query_length, key_length, out_dtype)

    def forward(self, q, k=None, v=None):
        query_length = q.shape[1]
        key_length = k.shape[1] if k is not None else query_length
        bias = self.compute_bias(query_length, key_length, device=q.device, dtype=q.dtype)   # contiguous by construction
        return q, k, v, bias
