"""Oracle for the RMS-scaled AdamW step with optional Kahan compensation (TEST INFRASTRUCTURE, not product).

Restates /root/reference/src/utils/adamw_scaled.py:154-211 (`AdamWScale._adamwscaled`, the per-tensor path; the
`_foreach` path :213-281 is the same arithmetic regrouped) with torch CPU tensors used as an ndarray library:

    m  <- beta1 m + (1 - beta1) g                                   :169
    v  <- beta2 v + (1 - beta2) g^2                                 :170
    denom = sqrt(v) + eps                                           :171
    step  = lr [* sqrt(1 - beta2^t) / (1 - beta1^t)]                :173-177
    step  = step * max(1e-3, rms(p)),  rms(p) = ||p||_2 / sqrt(numel)   :180   (p BEFORE the update)
    Kahan (16-bit p):  c <- c - step m / denom;  t = p;  p <- p + c;  c <- c + (t - p)     :182-192
    otherwise:         p <- p - step m / denom                                              :194
    p <- p - lr wd p                                                                        :204-205

`step_exact` does this in fp64 (the mathematical answer); `step_like_reference` keeps every tensor in its own dtype and
rounds where the reference's in-place ops round (16-bit states and parameters round after every op, and the rms / step
size of a 16-bit parameter is itself a 16-bit tensor), which is what the CUDA kernel reproduces.

Parity pin: oracle/make_golden.py runs the reference optimizer itself on CPU (both paths) and freezes
tests/golden/adamw_*.npz; tests/test_adamw_scaled.py checks this file against them.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch


def bias_correction(step: int, beta1: float, beta2: float, correct_bias: bool = True) -> float:
    if not correct_bias:
        return 1.0
    return math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)


def reference_step_size(step: int, lr: float, beta1: float, beta2: float, correct_bias: bool = True):
    """lr [* sqrt(1 - beta2^t) / (1 - beta1^t)] with the reference's types (:173-177): the step counter is an int32
    TENSOR there, so beta ** step and the bias corrections are fp32 0-dim tensors and the quotient is an fp32 tensor;
    without bias correction the step size stays the python float lr.  (The type matters: an fp32 tensor times the rms
    of a bf16 parameter is fp32, the float lr times the same rms is bf16.)"""
    if not correct_bias:
        return lr
    t = torch.tensor(step, dtype=torch.int32)
    bias_correction1 = 1.0 - beta1 ** t
    bias_correction2 = 1.0 - beta2 ** t
    return lr * math.sqrt(bias_correction2) / bias_correction1


def step_exact(p, g, m, v, comp, step: int, lr: float, beta1: float, beta2: float, eps: float, weight_decay: float,
               correct_bias: bool = True):
    """fp64, no intermediate rounding.  comp may be None.  Returns (p, m, v, comp)."""
    p, g, m, v = (t.double() for t in (p, g, m, v))
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    denom = v.sqrt() + eps
    rms = float(p.norm(2)) / math.sqrt(p.numel())
    ss = lr * bias_correction(step, beta1, beta2, correct_bias) * max(1e-3, rms)
    upd = -ss * m / denom
    if comp is not None:
        comp = comp.double() + upd
        p_new = p + comp
        comp = comp + (p - p_new)          # exactly 0 in exact arithmetic
        p = p_new
    else:
        p = p + upd
    if weight_decay > 0.0:
        p = p + p * (-lr * weight_decay)
    return p, m, v, comp


def step_like_reference(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, comp: Optional[torch.Tensor],
                        step: int, lr: float, beta1: float, beta2: float, eps: float, weight_decay: float,
                        correct_bias: bool = True) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """Every tensor keeps its dtype; each line rounds where the reference's in-place op rounds.  Inputs are not modified."""
    p, m, v = p.clone(), m.clone(), v.clone()
    comp = comp.clone() if comp is not None else None
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = v.sqrt().add_(eps)
    step_size = reference_step_size(step, lr, beta1, beta2, correct_bias)
    rms = p.norm(2) / (p.numel() ** 0.5)                     # a 0-dim tensor of p's dtype
    step_size = step_size * max(1e-3, rms)                   # python max: the rms tensor when it exceeds 1e-3, else the float
    step_size = float(step_size)                             # (addcdiv_ takes the value through .item())
    if comp is not None:
        comp.addcdiv_(m, denom, value=-step_size)
        tmp = p.clone()
        p.add_(comp)
        tmp.sub_(p)
        comp.add_(tmp)
    else:
        p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)
    return p, m, v, comp
