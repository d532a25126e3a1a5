"""RMSNorm -- B200-native drop-in for /root/reference/src/model/ops/rms_norm.py.

    fast_rms_layernorm(X, W, eps)                      (reference :285-287)
    Fast_RMS_Layernorm.apply(X, W, eps)                (reference :250-283)
    torch.ops.b200t5.rmsnorm_fwd(X, weight, eps) -> (Y, rstd)            (reference :134-174)
    torch.ops.b200t5.rmsnorm_bwd(dy, x, weight, rstd, eps) -> (dx, dw)   (reference :186-236)

Hand-written sm_100a CUDA behind the C ABI (include/b200t5.h); no Triton, no fallback.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _cabi

__all__ = ["fast_rms_layernorm", "Fast_RMS_Layernorm", "rmsnorm_fwd", "rmsnorm_bwd"]


@torch.library.custom_op("b200t5::rmsnorm_fwd", mutates_args=(), device_types="cuda")
def rmsnorm_fwd(X: torch.Tensor, weight: torch.Tensor, eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    _cabi.require_cuda(X, weight)
    M, N = X.shape
    assert X.stride(-1) == 1                              # reference :143
    assert weight.shape == (N,) and weight.stride(-1) == 1
    assert N <= 65536 // X.element_size()                 # reference :155-157
    lib = _cabi.load()
    Y = torch.empty_like(X)
    assert Y.stride(-1) == 1
    rstd = torch.empty((M,), dtype=torch.float32, device=X.device)
    rc = lib.b200t5_rmsnorm_fwd(X.data_ptr(), weight.data_ptr(), Y.data_ptr(), rstd.data_ptr(), M, N,
                                X.stride(0) if M > 1 else N, Y.stride(0) if M > 1 else N, float(eps),
                                _cabi.dtype_code(X.dtype), _cabi.dtype_code(weight.dtype),
                                X.device.index, _cabi.stream_ptr(X.device))
    _cabi.check(rc, "b200t5_rmsnorm_fwd")
    return Y, rstd


@torch.library.register_fake("b200t5::rmsnorm_fwd")
def _rmsnorm_fwd_fake(X, weight, eps):
    M, N = X.shape
    return torch.empty_like(X), torch.empty((M,), dtype=torch.float32, device=X.device)


@torch.library.custom_op("b200t5::rmsnorm_bwd", mutates_args=(), device_types="cuda")
def rmsnorm_bwd(dy: torch.Tensor, x: torch.Tensor, weight: torch.Tensor, rstd: torch.Tensor,
                eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    _cabi.require_cuda(dy, x, weight, rstd)
    M, N = x.shape
    assert x.stride(-1) == 1
    assert dy.shape == (M, N)
    if dy.stride(-1) != 1:
        dy = dy.contiguous()
    assert weight.shape == (N,) and weight.stride(-1) == 1
    lib = _cabi.load()
    dx = torch.empty_like(x)
    dw = torch.empty((N,), dtype=weight.dtype, device=weight.device)
    nbytes = lib.b200t5_rmsnorm_bwd_workspace_bytes(N)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    rc = lib.b200t5_rmsnorm_bwd(dy.data_ptr(), x.data_ptr(), weight.data_ptr(), rstd.data_ptr(), dx.data_ptr(),
                                dw.data_ptr(), ws.data_ptr(), nbytes, M, N,
                                dy.stride(0) if M > 1 else N, x.stride(0) if M > 1 else N,
                                dx.stride(0) if M > 1 else N,
                                _cabi.dtype_code(x.dtype), _cabi.dtype_code(weight.dtype),
                                x.device.index, _cabi.stream_ptr(x.device))
    _cabi.check(rc, "b200t5_rmsnorm_bwd")
    return dx, dw


@torch.library.register_fake("b200t5::rmsnorm_bwd")
def _rmsnorm_bwd_fake(dy, x, weight, rstd, eps):
    return torch.empty_like(x), torch.empty((x.shape[1],), dtype=weight.dtype, device=weight.device)


class Fast_RMS_Layernorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, W, eps=1e-6):
        X_orig_shape = X.shape
        X = X.reshape(-1, X.shape[-1])
        y, rstd = torch.ops.b200t5.rmsnorm_fwd(X, W, eps)
        y = y.reshape(X_orig_shape)
        # y is not saved: the backward recomputes x_hat from X and rstd (reference :261-262)
        ctx.save_for_backward(X, W, rstd)
        ctx.x_shape_og = X_orig_shape
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dY):
        X, weight, rstd = ctx.saved_tensors
        dY = dY.reshape(-1, dY.shape[-1])
        assert dY.shape == X.shape
        dx, dw = torch.ops.b200t5.rmsnorm_bwd(dY, X, weight, rstd, ctx.eps)
        return dx.reshape(ctx.x_shape_og), dw, None


def fast_rms_layernorm(X, W, eps):
    return Fast_RMS_Layernorm.apply(X, W, eps)
