"""FlashAttention-2 with additive (T5 relative-position) bias -- B200-native drop-in.

Same call surface as the reference operator
(/root/reference/src/model/ops/flash_attention_v2_bias.py:228-288):

    flash_attention_v2_bias(q, k, v, bias, causal=False, sm_scale=None) -> o
    FlashAttentionAdditiveBias.apply(q, k, v, bias, causal, sm_scale)

and the same two-op structure underneath (reference :27-38, :91-106), minus the four Triton
tuning integers:

    torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, causal, sm_scale) -> (o, L)
    torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, bias, L, causal, sm_scale) -> (dq, dk, dv, ds)

The ops call hand-written sm_100a CUDA through the C ABI in include/b200t5.h.  There is no
Triton, no eager and no CPU fallback: without the built library or an sm_100 GPU they raise.

Differences from the reference, all deliberate:
  * dBias is summed over EVERY broadcast dim of bias (the reference only sums the batch dim and
    races on a broadcast head dim).  Up to 8 batch elements meet in one 16-bit TMA reduce-add
    group at L2; the groups (and a broadcast head dim) are then summed in fp32 and rounded once.
  * q/k/v/o whose strides are not multiples of 8 elements (or are 0: expanded views) are
    materialised with .contiguous() first -- TMA needs 16-byte aligned, non-degenerate strides --
    so `o` is then contiguous instead of sharing q's strides.  An expanded bias takes the
    strided pointer path in the forward and the repacked copy in the backward.
  * with bias=None the backward op returns an empty tensor in the `ds` slot (a custom op cannot
    return None); the autograd.Function turns it back into None.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch

from . import _cabi

__all__ = ["flash_attention_v2_bias", "FlashAttentionAdditiveBias", "attn_bias_fwd", "attn_bias_bwd", "attn_bias_bwd_f32dbias",
           "attn_bias_bwd_accum", "SharedBiasGrad", "flash_attention_v2_bias_shared"]


def _aligned(t: torch.Tensor) -> bool:
    """What the C ABI needs (TMA): unit last stride, 16-byte base, other strides multiples of 8."""
    if t.stride(-1) != 1 or t.data_ptr() % 16 != 0:
        return False
    # (stride 0 on a dimension > 1 -- an expanded view such as k.expand(...) for multi-query attention -- cannot be put in a
    #  TMA tensor map: such tensors are materialised by _prep; the reference Triton kernel honours them through its strides)
    return all(s % 8 == 0 and s > 0 for s, n in zip(t.stride()[:-1], t.shape[:-1]) if n > 1)


def _prep(t: torch.Tensor) -> torch.Tensor:
    return t if _aligned(t) else t.contiguous()


def _check_shapes(q, k, v, bias):
    if q.dim() != 4 or k.dim() != 4 or v.dim() != 4:
        raise ValueError("q, k, v must be (batch, heads, seq, head_dim)")
    B, H, M, D = q.shape
    N = k.shape[2]
    if k.shape != (B, H, N, D) or v.shape != (B, H, N, D):
        raise ValueError(f"shape mismatch: q {tuple(q.shape)} k {tuple(k.shape)} v {tuple(v.shape)}")
    if D not in (16, 32, 64, 128):                      # reference :233-234
        raise AssertionError(f"head dim {D} not in {{16, 32, 64, 128}}")
    if q.dtype not in (torch.float16, torch.bfloat16) or k.dtype != q.dtype or v.dtype != q.dtype:
        raise TypeError(f"q, k, v must share dtype float16 or bfloat16 (got {q.dtype}, {k.dtype}, {v.dtype})")
    if bias is not None:
        if bias.dim() != 4 or bias.shape[2] != M or bias.shape[3] != N or bias.shape[0] not in (1, B) \
                or bias.shape[1] not in (1, H):
            raise ValueError(f"bias shape {tuple(bias.shape)} is not broadcastable to {(B, H, M, N)}")
        if bias.dtype != q.dtype:
            raise TypeError(f"bias dtype {bias.dtype} must equal q dtype {q.dtype}")
    return B, H, M, N, D


def _base_params(q, k, v, bias, causal, sm_scale) -> _cabi.AttnParams:
    B, H, M, D = q.shape
    N = k.shape[2]
    p = _cabi.AttnParams()
    p.B, p.H, p.M, p.N, p.D = B, H, M, N, D
    p.dtype = _cabi.dtype_code(q.dtype)
    p.causal = 1 if causal else 0
    p.sm_scale = float(sm_scale)
    p.device = q.device.index if q.device.index is not None else torch.cuda.current_device()
    p.stream = _cabi.stream_ptr(q.device)
    # torch.use_deterministic_algorithms(True): fixed-order dQ / dBias accumulation in the backward (larger workspace)
    p.flags = _cabi.ATTN_DETERMINISTIC if torch.are_deterministic_algorithms_enabled() else 0
    p.q, p.q_strides = q.data_ptr(), _cabi.strides4(q)
    p.k, p.k_strides = k.data_ptr(), _cabi.strides4(k)
    p.v, p.v_strides = v.data_ptr(), _cabi.strides4(v)
    if bias is not None:
        p.bias, p.bias_strides = bias.data_ptr(), _cabi.strides4(bias)
        p.bias_B, p.bias_H = bias.shape[0], bias.shape[1]
    else:
        p.bias = None
        p.bias_B = p.bias_H = 1
    return p


@torch.library.custom_op("b200t5::attn_bias_fwd", mutates_args=(), device_types="cuda")
def attn_bias_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, bias: Optional[torch.Tensor],
                  causal: bool, sm_scale: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """(o, L): o like q (same strides when q is TMA-addressable), L = logsumexp rows, (B,H,M) fp32.
    Replaces flasht5::flash_attn_v2_fwd (reference :27-80)."""
    _cabi.require_cuda(q, k, v, bias)
    B, H, M, N, D = _check_shapes(q, k, v, bias)
    lib = _cabi.load()
    q, k, v = _prep(q), _prep(k), _prep(v)
    o = torch.empty_like(q)                               # reference :58 (keeps q's strides)
    L = torch.empty((B, H, M), device=q.device, dtype=torch.float32)
    p = _base_params(q, k, v, bias, causal, sm_scale)
    p.o, p.o_strides = o.data_ptr(), _cabi.strides4(o)
    p.lse = L.data_ptr()
    nbytes = lib.b200t5_attn_fwd_workspace_bytes(C.byref(p))        # > 0 only for bias rows a tensor map cannot address
    if nbytes:
        ws = torch.empty(nbytes, device=q.device, dtype=torch.uint8)
        p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
    _cabi.check(lib.b200t5_attn_fwd(C.byref(p)), "b200t5_attn_fwd")
    return o, L


@torch.library.register_fake("b200t5::attn_bias_fwd")
def _attn_bias_fwd_fake(q, k, v, bias, causal, sm_scale):
    B, H, M, D = q.shape
    return torch.empty_like(q), torch.empty((B, H, M), dtype=torch.float32, device=q.device)


@torch.library.custom_op("b200t5::attn_bias_bwd", mutates_args=(), device_types="cuda")
def attn_bias_bwd(o: torch.Tensor, do: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                  bias: Optional[torch.Tensor], L: torch.Tensor, causal: bool,
                  sm_scale: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(dq, dk, dv, ds): ds has bias's shape (empty tensor when bias is None).
    Replaces flasht5::flash_attn_v2_bwd (reference :91-217) incl. the trailing ds.sum(0)."""
    _cabi.require_cuda(o, do, q, k, v, bias, L)
    B, H, M, N, D = _check_shapes(q, k, v, bias)
    lib = _cabi.load()
    q, k, v, o, do = _prep(q), _prep(k), _prep(v), _prep(o), _prep(do)
    L = L.contiguous()
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ds = torch.empty_like(bias, memory_format=torch.contiguous_format) if bias is not None \
        else torch.empty(0, dtype=q.dtype, device=q.device)
    p = _base_params(q, k, v, bias, causal, sm_scale)
    p.o, p.o_strides = o.data_ptr(), _cabi.strides4(o)
    p.lse = L.data_ptr()
    p.dout, p.do_strides = do.data_ptr(), _cabi.strides4(do)
    p.dq, p.dq_strides = dq.data_ptr(), _cabi.strides4(dq)
    p.dk, p.dk_strides = dk.data_ptr(), _cabi.strides4(dk)
    p.dv, p.dv_strides = dv.data_ptr(), _cabi.strides4(dv)
    if bias is not None:
        p.dbias, p.dbias_strides = ds.data_ptr(), _cabi.strides4(ds)
    nbytes = lib.b200t5_attn_bwd_workspace_bytes(C.byref(p))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)   # caching allocator: 512-byte aligned
    p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
    _cabi.check(lib.b200t5_attn_bwd(C.byref(p)), "b200t5_attn_bwd")
    return dq, dk, dv, ds


@torch.library.custom_op("b200t5::attn_bias_bwd_f32dbias", mutates_args=(), device_types="cuda")
def attn_bias_bwd_f32dbias(o: torch.Tensor, do: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                           bias: torch.Tensor, L: torch.Tensor, causal: bool,
                           sm_scale: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """As attn_bias_bwd, but dBias comes back as an UNROUNDED fp32 tensor (C ABI flag B200T5_ATTN_DBIAS_F32; head dims
    16 / 32 / 64): the data-parallel exchange sums it across ranks and rounds once (data_parallel.allreduce_dbias_f32)."""
    _cabi.require_cuda(o, do, q, k, v, bias, L)
    B, H, M, N, D = _check_shapes(q, k, v, bias)
    lib = _cabi.load()
    q, k, v, o, do = _prep(q), _prep(k), _prep(v), _prep(o), _prep(do)
    L = L.contiguous()
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ds = torch.empty(bias.shape, dtype=torch.float32, device=q.device)
    p = _base_params(q, k, v, bias, causal, sm_scale)
    p.flags |= _cabi.ATTN_DBIAS_F32
    p.o, p.o_strides = o.data_ptr(), _cabi.strides4(o)
    p.lse = L.data_ptr()
    p.dout, p.do_strides = do.data_ptr(), _cabi.strides4(do)
    p.dq, p.dq_strides = dq.data_ptr(), _cabi.strides4(dq)
    p.dk, p.dk_strides = dk.data_ptr(), _cabi.strides4(dk)
    p.dv, p.dv_strides = dv.data_ptr(), _cabi.strides4(dv)
    p.dbias, p.dbias_strides = ds.data_ptr(), _cabi.strides4(ds)
    nbytes = lib.b200t5_attn_bwd_workspace_bytes(C.byref(p))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
    _cabi.check(lib.b200t5_attn_bwd(C.byref(p)), "b200t5_attn_bwd")
    return dq, dk, dv, ds


@torch.library.register_fake("b200t5::attn_bias_bwd_f32dbias")
def _attn_bias_bwd_f32dbias_fake(o, do, q, k, v, bias, L, causal, sm_scale):
    return torch.empty_like(q), torch.empty_like(k), torch.empty_like(v), torch.empty(bias.shape, dtype=torch.float32, device=q.device)


@torch.library.custom_op("b200t5::attn_bias_bwd_accum", mutates_args=("dbias_acc",), device_types="cuda")
def attn_bias_bwd_accum(o: torch.Tensor, do: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                        bias: torch.Tensor, L: torch.Tensor, dbias_acc: torch.Tensor, causal: bool,
                        sm_scale: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(dq, dk, dv); the batch-summed dS is ADDED, unrounded, into the fp32 tensor `dbias_acc` of bias's shape (C ABI flags
    B200T5_ATTN_DBIAS_F32 | B200T5_ATTN_DBIAS_ACCUMULATE; head dims 16 / 32 / 64).  SURVEY.md section 8 row f2: the layers of a
    stack share one bias (reference modeling_flash_t5.py:452-455) and accumulate its gradient here instead of in autograd."""
    _cabi.require_cuda(o, do, q, k, v, bias, L, dbias_acc)
    B, H, M, N, D = _check_shapes(q, k, v, bias)
    if dbias_acc.dtype != torch.float32 or dbias_acc.shape != bias.shape:
        raise ValueError("dbias_acc must be an fp32 tensor of bias's shape")
    lib = _cabi.load()
    q, k, v, o, do = _prep(q), _prep(k), _prep(v), _prep(o), _prep(do)
    L = L.contiguous()
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    p = _base_params(q, k, v, bias, causal, sm_scale)
    p.flags |= _cabi.ATTN_DBIAS_F32 | _cabi.ATTN_DBIAS_ACCUMULATE
    p.o, p.o_strides = o.data_ptr(), _cabi.strides4(o)
    p.lse = L.data_ptr()
    p.dout, p.do_strides = do.data_ptr(), _cabi.strides4(do)
    p.dq, p.dq_strides = dq.data_ptr(), _cabi.strides4(dq)
    p.dk, p.dk_strides = dk.data_ptr(), _cabi.strides4(dk)
    p.dv, p.dv_strides = dv.data_ptr(), _cabi.strides4(dv)
    p.dbias, p.dbias_strides = dbias_acc.data_ptr(), _cabi.strides4(dbias_acc)
    nbytes = lib.b200t5_attn_bwd_workspace_bytes(C.byref(p))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
    _cabi.check(lib.b200t5_attn_bwd(C.byref(p)), "b200t5_attn_bwd")
    return dq, dk, dv


@torch.library.register_fake("b200t5::attn_bias_bwd_accum")
def _attn_bias_bwd_accum_fake(o, do, q, k, v, bias, L, dbias_acc, causal, sm_scale):
    return torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)


@torch.library.register_fake("b200t5::attn_bias_bwd")
def _attn_bias_bwd_fake(o, do, q, k, v, bias, L, causal, sm_scale):
    ds = torch.empty_like(bias, memory_format=torch.contiguous_format) if bias is not None \
        else torch.empty(0, dtype=q.dtype, device=q.device)
    return torch.empty_like(q), torch.empty_like(k), torch.empty_like(v), ds


class FlashAttentionAdditiveBias(torch.autograd.Function):
    """Mirror of the reference autograd.Function (reference :228-271)."""

    @staticmethod
    def forward(ctx, q, k, v, bias, causal, sm_scale):
        Dq, Dk, Dv = q.shape[-1], k.shape[-1], v.shape[-1]
        assert Dq == Dk == Dv
        assert Dk in {16, 32, 64, 128}
        if sm_scale is None:
            sm_scale = 1.0 / math.sqrt(Dq)               # reference :239-240
        o, L = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, bool(causal), float(sm_scale))
        ctx.save_for_backward(q, k, v, bias, o, L)
        ctx.sm_scale = float(sm_scale)
        ctx.causal = bool(causal)
        return o

    @staticmethod
    def backward(ctx, do, *ignored):
        q, k, v, bias, o, L = ctx.saved_tensors
        dq, dk, dv, ds = torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, bias, L, ctx.causal, ctx.sm_scale)
        return dq, dk, dv, (ds if bias is not None else None), None, None


class SharedBiasGrad:
    """Gradient accumulator of ONE bias tensor shared by several attention layers.

    The reference computes the position bias once per stack and hands the same tensor to every layer
    (modeling_flash_t5.py:437-457), so autograd adds L dense (1, H, M, N) gradients per stack, each already rounded to 16
    bits.  With one `SharedBiasGrad()` per (stack, step) passed to `flash_attention_v2_bias_shared`, every layer's backward
    adds its unrounded batch-summed dS into one persistent fp32 buffer inside the finalize kernel, and autograd receives a
    single gradient -- from the layer whose backward runs last -- rounded once."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None
        self.pending = 0


class FlashAttentionSharedBias(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, bias, acc, causal, sm_scale):
        assert q.shape[-1] == k.shape[-1] == v.shape[-1] and q.shape[-1] in {16, 32, 64, 128}
        if sm_scale is None:
            sm_scale = 1.0 / math.sqrt(q.shape[-1])
        o, L = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, bool(causal), float(sm_scale))
        ctx.save_for_backward(q, k, v, bias, o, L)
        ctx.sm_scale, ctx.causal, ctx.acc = float(sm_scale), bool(causal), acc
        acc.pending += 1
        return o

    @staticmethod
    def backward(ctx, do, *ignored):
        q, k, v, bias, o, L = ctx.saved_tensors
        acc = ctx.acc
        if acc.buf is None:
            acc.buf = torch.zeros(bias.shape, dtype=torch.float32, device=bias.device)
        if q.shape[-1] <= 64:
            dq, dk, dv = torch.ops.b200t5.attn_bias_bwd_accum(o, do, q, k, v, bias, L, acc.buf, ctx.causal, ctx.sm_scale)
        else:                                    # D = 128: the fp32 output is not implemented in that kernel; add in fp32 here
            dq, dk, dv, ds = torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, bias, L, ctx.causal, ctx.sm_scale)
            acc.buf += ds
        acc.pending -= 1
        g = None
        if acc.pending == 0:                     # the last backward of the stack hands the one gradient to autograd
            g, acc.buf = acc.buf.to(bias.dtype), None
        return dq, dk, dv, g, None, None, None


def flash_attention_v2_bias_shared(q, k, v, bias, acc: SharedBiasGrad, causal=False, sm_scale=None):
    """`flash_attention_v2_bias` for a bias shared by several layers: same result, the bias gradient is accumulated in fp32
    across the layers that were given the same `acc` and reaches autograd once (see SharedBiasGrad)."""
    if bias is None or not bias.requires_grad:
        return FlashAttentionAdditiveBias.apply(q, k, v, bias, causal, sm_scale)
    return FlashAttentionSharedBias.apply(q, k, v, bias, acc, causal, sm_scale)


def flash_attention_v2_bias(q, k, v, bias, causal=False, sm_scale=None):
    """softmax(q k^T * sm_scale + bias [+ causal mask]) v   with q:(B,H,M,D) k,v:(B,H,N,D)
    bias:(1|B, 1|H, M, N) or None.  Default sm_scale = 1/sqrt(D).  (reference :274-288)"""
    return FlashAttentionAdditiveBias.apply(q, k, v, bias, causal, sm_scale)
