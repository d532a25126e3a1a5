"""Cross-entropy + z-loss -- B200-native drop-in for
/root/reference/src/model/ops/cross_entropy_loss.py.

    cross_entropy_loss(logits, labels, precomputed_lse=None, label_smoothing=0.0, logit_scale=1.0,
                       lse_square_scale=0.0, ignore_index=-100, inplace_backward=False,
                       process_group=None) -> (losses, z_losses)            (reference :388-426)
    CrossEntropyLoss.apply(...)                                             (reference :280-385)
    torch.ops.b200t5.ce_fwd / torch.ops.b200t5.ce_bwd                       (reference :164-274)

The vocab-parallel branch (process_group is not None, reference :324-351) is dead code in the
reference (no caller passes a group) and is out of scope here: it raises NotImplementedError.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _cabi

__all__ = ["cross_entropy_loss", "CrossEntropyLoss", "ce_fwd", "ce_bwd"]


@torch.library.custom_op("b200t5::ce_fwd", mutates_args=(), device_types="cuda")
def ce_fwd(logits: torch.Tensor, labels: torch.Tensor, precomputed_lse: Optional[torch.Tensor],
           smoothing: float, logit_scale: float, lse_square_scale: float,
           ignore_index: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(losses, z_losses, lse), all (rows,) fp32."""
    _cabi.require_cuda(logits, labels, precomputed_lse)
    if logits.stride(-1) != 1:
        logits = logits.contiguous()                       # reference :182-183
    n_rows, n_cols = logits.shape
    assert labels.shape == (n_rows,)
    labels = labels.to(torch.int64).contiguous()
    lib = _cabi.load()
    losses = torch.empty(n_rows, dtype=torch.float32, device=logits.device)
    z_losses = torch.empty(n_rows, dtype=torch.float32, device=logits.device)
    use_pre = precomputed_lse is not None
    if use_pre:
        assert precomputed_lse.shape == (n_rows,)
        lse = precomputed_lse.to(torch.float32).contiguous().clone()
    else:
        lse = torch.empty(n_rows, dtype=torch.float32, device=logits.device)
    rc = lib.b200t5_ce_fwd(logits.data_ptr(), labels.data_ptr(), losses.data_ptr(), z_losses.data_ptr(),
                           lse.data_ptr(), 1 if use_pre else 0, n_rows, n_cols,
                           logits.stride(0) if n_rows > 1 else n_cols,
                           float(smoothing), float(logit_scale), float(lse_square_scale), int(ignore_index),
                           _cabi.dtype_code(logits.dtype), logits.device.index, _cabi.stream_ptr(logits.device))
    _cabi.check(rc, "b200t5_ce_fwd")
    return losses, z_losses, lse


@torch.library.register_fake("b200t5::ce_fwd")
def _ce_fwd_fake(logits, labels, precomputed_lse, smoothing, logit_scale, lse_square_scale, ignore_index):
    n = logits.shape[0]
    mk = lambda: torch.empty(n, dtype=torch.float32, device=logits.device)   # noqa: E731
    return mk(), mk(), mk()


def _ce_bwd_impl(dlosses, logits, lse, labels, out, smoothing, logit_scale, lse_square_scale, ignore_index):
    lib = _cabi.load()
    n_rows, n_cols = logits.shape
    labels = labels.to(torch.int64).contiguous()
    dlosses = dlosses.to(torch.float32)
    rc = lib.b200t5_ce_bwd(logits.data_ptr(), labels.data_ptr(), lse.data_ptr(), dlosses.data_ptr(),
                           dlosses.stride(0) if n_rows > 1 else 1, out.data_ptr(), n_rows, n_cols,
                           logits.stride(0) if n_rows > 1 else n_cols, out.stride(0) if n_rows > 1 else n_cols,
                           float(smoothing), float(logit_scale), float(lse_square_scale), int(ignore_index),
                           _cabi.dtype_code(logits.dtype), logits.device.index, _cabi.stream_ptr(logits.device))
    _cabi.check(rc, "b200t5_ce_bwd")


@torch.library.custom_op("b200t5::ce_bwd", mutates_args=(), device_types="cuda")
def ce_bwd(dlosses: torch.Tensor, logits: torch.Tensor, lse: torch.Tensor, labels: torch.Tensor,
           smoothing: float, logit_scale: float, lse_square_scale: float, ignore_index: int) -> torch.Tensor:
    _cabi.require_cuda(dlosses, logits, lse, labels)
    if logits.stride(-1) != 1:
        logits = logits.contiguous()
    dlogits = torch.empty_like(logits)
    _ce_bwd_impl(dlosses, logits, lse, labels, dlogits, smoothing, logit_scale, lse_square_scale, ignore_index)
    return dlogits


@torch.library.register_fake("b200t5::ce_bwd")
def _ce_bwd_fake(dlosses, logits, lse, labels, smoothing, logit_scale, lse_square_scale, ignore_index):
    return torch.empty_like(logits)


@torch.library.custom_op("b200t5::ce_bwd_inplace", mutates_args={"logits"}, device_types="cuda")
def ce_bwd_inplace(dlosses: torch.Tensor, logits: torch.Tensor, lse: torch.Tensor, labels: torch.Tensor,
                   smoothing: float, logit_scale: float, lse_square_scale: float, ignore_index: int) -> None:
    """dlogits overwrite `logits` (reference inplace_backward, :247,274,382-383)."""
    _cabi.require_cuda(dlosses, logits, lse, labels)
    assert logits.stride(-1) == 1
    _ce_bwd_impl(dlosses, logits, lse, labels, logits, smoothing, logit_scale, lse_square_scale, ignore_index)


class CrossEntropyLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, precomputed_lse=None, smoothing=0.0, logit_scale=1.0, lse_square_scale=0.0,
                ignore_index=-100, inplace_backward=False, process_group=None):
        if process_group is not None:
            raise NotImplementedError("vocab-parallel cross-entropy (process_group) is out of scope; "
                                      "the reference never passes a group (modeling_flash_t5.py:64-68)")
        n_rows, n_cols = logits.shape
        assert labels.shape == (n_rows,)
        use_precomputed_lse = precomputed_lse is not None and logit_scale == 1.0 and smoothing == 0.0   # :307
        if inplace_backward and logits.stride(-1) != 1:
            logits = logits.contiguous()
        losses, z_losses, lse = torch.ops.b200t5.ce_fwd(
            logits, labels, precomputed_lse if use_precomputed_lse else None, smoothing, logit_scale,
            lse_square_scale, ignore_index)
        ctx.save_for_backward(logits, lse, labels)
        ctx.mark_non_differentiable(z_losses)
        ctx.smoothing = smoothing
        ctx.logit_scale = logit_scale
        ctx.lse_square_scale = lse_square_scale
        ctx.ignore_index = ignore_index
        ctx.inplace_backward = inplace_backward
        return losses, z_losses

    @staticmethod
    def backward(ctx, grad_losses, grad_z_losses):
        del grad_z_losses                                   # z_losses are only for logging
        logits, lse, labels = ctx.saved_tensors
        args = (ctx.smoothing, ctx.logit_scale, ctx.lse_square_scale, ctx.ignore_index)
        if ctx.inplace_backward:
            torch.ops.b200t5.ce_bwd_inplace(grad_losses, logits, lse, labels, *args)
            dlogits = logits
        else:
            dlogits = torch.ops.b200t5.ce_bwd(grad_losses, logits, lse, labels, *args)
        return dlogits, None, None, None, None, None, None, None, None


def cross_entropy_loss(logits: torch.Tensor, labels: torch.Tensor, precomputed_lse: Optional[torch.Tensor] = None,
                       label_smoothing: float = 0.0, logit_scale: float = 1.0, lse_square_scale: float = 0.0,
                       ignore_index=-100, inplace_backward: bool = False,
                       process_group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """logits (rows, vocab), labels (rows,) -> (losses, z_losses), both (rows,) fp32; ignored rows give 0.
    loss = lse - x[label] (+ smoothing term) + lse_square_scale * lse^2."""
    return CrossEntropyLoss.apply(logits.view(-1, logits.shape[-1]), labels.view(-1), precomputed_lse,
                                  label_smoothing, logit_scale, lse_square_scale, ignore_index, inplace_backward,
                                  process_group)
