"""FlashAttention-2 with the T5 relative-position bias computed INSIDE the kernels -- the B200-native operator for
the reference's `attention_type == "fa2_rpe"` call (/root/reference/src/model/modeling_flash_t5.py:275-279):

    flash_attn_func(q, k, v, softmax_scale=..., causal=...,
                    rpe_weights=pe_encoding.relative_attention_bias.weight.t(), rpe_max_distance=...)

Here:

    flash_attention_v2_rpe(q, k, v, rpe_weights, rpe_max_distance, causal=False, sm_scale=None) -> o

with q:(B,H,M,D), k,v:(B,H,N,D) (the layout of `flash_attention_v2_bias`, i.e. the model's (B,S,H,D) memory
permuted) and rpe_weights:(H, num_buckets) -- the transposed embedding weight, exactly what the reference passes.
The result equals `flash_attention_v2_bias(q, k, v, bias)` with
bias = RelativePositionalEncoding(num_buckets, rpe_max_distance, H, ., bidirectional=not causal).compute_bias(M, N)
cast to q.dtype (/root/reference/src/utils/positional_encoding.py:73-102), and the gradient that reaches
`rpe_weights` equals what autograd would scatter back through that gather -- but no (1,H,M,N) bias tensor is ever
built and no (1,H,M,N) gradient is handed back to autograd.

How: the bucket of every relative position is evaluated once with the reference formula into a lookup table (as
positional_encoding.py here does), the positions beyond which the table is constant (|n - m| >= max_distance for
T5) are read off it, and a small per-head "band" of bias values over the remaining relative positions is built by
one tiny kernel.  The attention kernels keep that band in shared memory: a 128x128 tile entirely beyond the last
distinct bucket adds one scalar, the others index band[n - m].  The flash-attention fork that implements the
reference's surface is not part of the reference checkout ("parity unpinned" for that arm);
parity here is pinned on the reference's own dense composition (tests/golden/rpe_*.npz).

When the table has more than ~7.6k distinct relative positions between its constant ends (not a T5 table) the
call composes the dense producer + dense-bias kernels instead (still CUDA, still this library).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Tuple

import torch

from . import _cabi
from .flash_attention_v2_bias import _base_params, _check_shapes, _prep, flash_attention_v2_bias
from .positional_encoding import RelativePositionalEncoding, _T5Bias

__all__ = ["flash_attention_v2_rpe", "FlashAttentionRPE", "rpe_band", "attn_rpe_fwd", "attn_rpe_bwd",
           "bucket_lut", "constant_ends", "band_len", "fused_default", "BAND_PAD", "MAX_BAND_LEN"]

BAND_PAD = 255          # kernels.h: kRpeBandPad
MAX_BAND_LEN = 4096     # kernels.h: kRpeMaxBandLen

_LUT_CACHE = {}          # (M, N, num_buckets, max_distance, bidirectional, device) -> (lut, lut_zero, const_lo, const_hi)
_LUT_CACHE_MAX = 64      # shapes a process meets are few; bound it anyway (oldest entry goes first)


def constant_ends(lut_cpu: torch.Tensor, lut_zero: int) -> Tuple[int, int]:
    """(const_lo, const_hi): every rel <= const_lo maps to lut[const_lo], every rel >= const_hi to lut[const_hi],
    const_lo < const_hi.  Host logic (runs on a CPU copy of the lookup table, once per shape)."""
    n = lut_cpu.numel()
    assert n >= 1
    if n == 1:
        return -lut_zero - 1, -lut_zero
    first, last = int(lut_cpu[0]), int(lut_cpu[-1])
    neq_first = (lut_cpu != first).nonzero()
    i_lo = int(neq_first[0]) - 1 if neq_first.numel() else n - 1          # last index of the leading constant run
    neq_last = (lut_cpu != last).nonzero()
    i_hi = int(neq_last[-1]) + 1 if neq_last.numel() else 0               # first index of the trailing constant run
    if i_lo >= i_hi:                                                       # the whole table is one bucket
        i_lo = i_hi - 1
    return i_lo - lut_zero, i_hi - lut_zero


@torch.compiler.disable      # host-side table building (one device->host copy per new shape): not something to trace
def bucket_lut(M: int, N: int, num_buckets: int, max_distance: int, bidirectional: bool, device):
    """(lut int32 on `device`, lut_zero, const_lo, const_hi) for relative positions -(M-1) .. N-1, cached.
    The buckets come from the reference formula evaluated on `device` (positional_encoding.py)."""
    key = (M, N, num_buckets, max_distance, bool(bidirectional), str(device))
    hit = _LUT_CACHE.get(key)
    if hit is None:
        rel = torch.arange(-(M - 1), N, dtype=torch.long, device=device)
        lut = RelativePositionalEncoding._relative_position_bucket(
            rel, bidirectional=bidirectional, num_buckets=num_buckets, max_distance=max_distance).to(torch.int32)
        lo, hi = constant_ends(lut.cpu(), M - 1)
        hit = (lut, M - 1, lo, hi)
        if len(_LUT_CACHE) >= _LUT_CACHE_MAX:
            _LUT_CACHE.pop(next(iter(_LUT_CACHE)))
        _LUT_CACHE[key] = hit
    return hit


def band_len(const_lo: int, const_hi: int) -> int:
    return const_hi - const_lo + 2 * BAND_PAD + 1


def _rpe_struct(table, lut, lut_zero, const_lo, const_hi, band, dtable=None) -> _cabi.RpeParams:
    r = _cabi.RpeParams()
    if table is not None:
        r.table = table.data_ptr()
        r.table_stride_b, r.table_stride_h = table.stride(0), table.stride(1)
        r.table_dtype = _cabi.dtype_code(table.dtype)
        r.num_buckets = table.shape[0]
    if lut is not None:
        r.lut = lut.data_ptr()
        r.lut_zero, r.lut_len = lut_zero, lut.numel()
    r.const_lo, r.const_hi = const_lo, const_hi
    r.band = band.data_ptr()
    if dtable is not None:
        r.dtable = dtable.data_ptr()
        r.num_buckets = dtable.shape[0]
    return r


@torch.library.custom_op("b200t5::rpe_band", mutates_args=(), device_types="cuda")
def rpe_band(table: torch.Tensor, lut: torch.Tensor, lut_zero: int, const_lo: int, const_hi: int,
             io_dtype: torch.dtype) -> torch.Tensor:
    """table (num_buckets, H) any strides -> band (H, band_len) fp32: the bias of every relative position
    const_lo-255 .. const_hi+255 per head, rounded to `io_dtype` like the dense path's bias cast."""
    _cabi.require_cuda(table, lut)
    lib = _cabi.load()
    H = table.shape[1]
    band = torch.empty((H, band_len(const_lo, const_hi)), dtype=torch.float32, device=table.device)
    r = _rpe_struct(table, lut, lut_zero, const_lo, const_hi, band)
    rc = lib.b200t5_rpe_band(C.byref(r), H, _cabi.dtype_code(io_dtype), table.device.index,
                             _cabi.stream_ptr(table.device))
    _cabi.check(rc, "b200t5_rpe_band")
    return band


@torch.library.register_fake("b200t5::rpe_band")
def _rpe_band_fake(table, lut, lut_zero, const_lo, const_hi, io_dtype):
    return torch.empty((table.shape[1], band_len(const_lo, const_hi)), dtype=torch.float32, device=table.device)


@torch.library.custom_op("b200t5::attn_rpe_fwd", mutates_args=(), device_types="cuda")
def attn_rpe_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, band: torch.Tensor, const_lo: int,
                 const_hi: int, causal: bool, sm_scale: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """(o, L) like b200t5::attn_bias_fwd, the bias coming from `band` (see rpe_band)."""
    _cabi.require_cuda(q, k, v, band)
    B, H, M, N, D = _check_shapes(q, k, v, None)
    if band.shape != (H, band_len(const_lo, const_hi)) or band.dtype != torch.float32 or not band.is_contiguous():
        raise ValueError("band must be the contiguous fp32 (H, band_len) tensor made by rpe_band")
    lib = _cabi.load()
    q, k, v = _prep(q), _prep(k), _prep(v)
    o = torch.empty_like(q)
    L = torch.empty((B, H, M), device=q.device, dtype=torch.float32)
    p = _base_params(q, k, v, None, causal, sm_scale)
    p.o, p.o_strides = o.data_ptr(), _cabi.strides4(o)
    p.lse = L.data_ptr()
    r = _rpe_struct(None, None, 0, const_lo, const_hi, band)
    _cabi.check(lib.b200t5_attn_rpe_fwd(C.byref(p), C.byref(r)), "b200t5_attn_rpe_fwd")
    return o, L


@torch.library.register_fake("b200t5::attn_rpe_fwd")
def _attn_rpe_fwd_fake(q, k, v, band, const_lo, const_hi, causal, sm_scale):
    B, H, M, D = q.shape
    return torch.empty_like(q), torch.empty((B, H, M), dtype=torch.float32, device=q.device)


@torch.library.custom_op("b200t5::attn_rpe_bwd", mutates_args=(), device_types="cuda")
def attn_rpe_bwd(o: torch.Tensor, do: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                 band: torch.Tensor, lut: torch.Tensor, lut_zero: int, const_lo: int, const_hi: int,
                 num_buckets: int, L: torch.Tensor, causal: bool,
                 sm_scale: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(dq, dk, dv, dtable): dtable (num_buckets, H) fp32 = the gradient of the embedding table."""
    _cabi.require_cuda(o, do, q, k, v, band, lut, L)
    B, H, M, N, D = _check_shapes(q, k, v, None)
    lib = _cabi.load()
    q, k, v, o, do = _prep(q), _prep(k), _prep(v), _prep(o), _prep(do)
    L = L.contiguous()
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    dtable = torch.empty((num_buckets, H), dtype=torch.float32, device=q.device)
    p = _base_params(q, k, v, None, causal, sm_scale)
    p.o, p.o_strides = o.data_ptr(), _cabi.strides4(o)
    p.lse = L.data_ptr()
    p.dout, p.do_strides = do.data_ptr(), _cabi.strides4(do)
    p.dq, p.dq_strides = dq.data_ptr(), _cabi.strides4(dq)
    p.dk, p.dk_strides = dk.data_ptr(), _cabi.strides4(dk)
    p.dv, p.dv_strides = dv.data_ptr(), _cabi.strides4(dv)
    r = _rpe_struct(None, lut, lut_zero, const_lo, const_hi, band, dtable)
    nbytes = lib.b200t5_attn_rpe_bwd_workspace_bytes(C.byref(p), C.byref(r))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
    _cabi.check(lib.b200t5_attn_rpe_bwd(C.byref(p), C.byref(r)), "b200t5_attn_rpe_bwd")
    return dq, dk, dv, dtable


@torch.library.register_fake("b200t5::attn_rpe_bwd")
def _attn_rpe_bwd_fake(o, do, q, k, v, band, lut, lut_zero, const_lo, const_hi, num_buckets, L, causal, sm_scale):
    return (torch.empty_like(q), torch.empty_like(k), torch.empty_like(v),
            torch.empty((num_buckets, q.shape[1]), dtype=torch.float32, device=q.device))


class FlashAttentionRPE(torch.autograd.Function):
    """q, k, v, table (num_buckets, H) -> o; backward returns dq, dk, dv and the table gradient."""

    @staticmethod
    def forward(ctx, q, k, v, table, lut, lut_zero, const_lo, const_hi, causal, sm_scale):
        band = torch.ops.b200t5.rpe_band(table, lut, lut_zero, const_lo, const_hi, q.dtype)
        o, L = torch.ops.b200t5.attn_rpe_fwd(q, k, v, band, const_lo, const_hi, bool(causal), float(sm_scale))
        ctx.save_for_backward(q, k, v, o, L, band, lut)
        ctx.args = (lut_zero, const_lo, const_hi, table.shape[0], bool(causal), float(sm_scale), table.dtype)
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, L, band, lut = ctx.saved_tensors
        lut_zero, const_lo, const_hi, num_buckets, causal, sm_scale, table_dtype = ctx.args
        dq, dk, dv, dtable = torch.ops.b200t5.attn_rpe_bwd(o, do, q, k, v, band, lut, lut_zero, const_lo, const_hi,
                                                           num_buckets, L, causal, sm_scale)
        return dq, dk, dv, dtable.to(table_dtype), None, None, None, None, None, None


def fused_default() -> bool:
    """Whether flash_attention_v2_rpe takes the in-kernel path when `fused` is not given: yes, unless
    B200T5_RPE_FUSED=0 (developer switch for A/B runs against the composition of the dense kernels)."""
    return os.environ.get("B200T5_RPE_FUSED", "1") != "0"


def flash_attention_v2_rpe(q, k, v, rpe_weights, rpe_max_distance, causal=False, sm_scale=None,
                           bidirectional: Optional[bool] = None, fused: Optional[bool] = None):
    """softmax(q k^T * sm_scale + T5 bias(rpe_weights) [+ causal mask]) v.

    rpe_weights: (H, num_buckets) -- `relative_attention_bias.weight.t()` as the reference passes it (any strides).
    bidirectional defaults to `not causal` (the reference builds the encoder's encoding bidirectional and the
    decoder's unidirectional, modeling_flash_t5.py:208-212).  Default sm_scale = 1/sqrt(D).
    fused: True = bias computed inside the attention kernels; False = dense producer kernel + dense-bias kernels
    (same results: the in-kernel band holds the same 16-bit-rounded values); None = fused_default()."""
    if rpe_weights.dim() != 2 or rpe_weights.shape[0] != q.shape[1]:
        raise ValueError(f"rpe_weights must be (H, num_buckets) with H = {q.shape[1]} (got {tuple(rpe_weights.shape)})")
    D = q.shape[-1]
    assert q.shape[-1] == k.shape[-1] == v.shape[-1]
    assert D in {16, 32, 64, 128}
    if sm_scale is None:
        sm_scale = 1.0 / math.sqrt(D)
    if bidirectional is None:
        bidirectional = not causal
    _cabi.require_cuda(q, k, v, rpe_weights)
    M, N = q.shape[2], k.shape[2]
    table = rpe_weights.t()                                   # (num_buckets, H) view; autograd transposes the grad back
    lut, lut_zero, const_lo, const_hi = bucket_lut(M, N, table.shape[0], int(rpe_max_distance), bidirectional,
                                                   q.device)
    if fused is None:
        fused = fused_default()
    if not fused or band_len(const_lo, const_hi) > MAX_BAND_LEN:
        # composition of the dense kernels (also the route for a table that is not T5-shaped): materialise the
        # bias with the producer kernel, run the dense-bias attention, scatter dBias back with the producer backward
        bias = _T5Bias.apply(table.contiguous(), lut, lut_zero, None, None, M, N, q.dtype)
        return flash_attention_v2_bias(q, k, v, bias, causal, sm_scale)
    return FlashAttentionRPE.apply(q, k, v, table, lut, lut_zero, const_lo, const_hi, causal, sm_scale)
