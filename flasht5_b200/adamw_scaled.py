"""AdamWScale -- B200-native drop-in for the reference optimizer (/root/reference/src/utils/adamw_scaled.py:10-281):
AdamW whose step is scaled by max(1e-3, rms(parameter)), with optional Kahan compensation for 16-bit parameters.

Same constructor, same state (`step`, `exp_avg`, `exp_avg_sq`, `kahan_comp`: checkpoints of the reference optimizer
load unchanged), same arithmetic including where the reference's in-place tensor ops round.  What changes underneath:
the reference's per-tensor path launches ~10 small kernels and one `.item()` synchronisation per parameter, its
`_foreach` path ~15 passes per dtype group plus a `.item()` per tensor (:247); here all tensors of a
(device, parameter dtype, state dtype, kahan) group are updated by three launches of hand-written sm_100a CUDA through
the C ABI (`b200t5_adamw_scale_step`, csrc/adamw.cu) with no host synchronisation.  The `foreach` argument is accepted
and ignored.  No CPU fallback: parameters must live on an sm_100 device.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, Tuple

import torch
from torch import nn
from torch.optim import Optimizer

from . import _cabi

__all__ = ["AdamWScale", "step_size_terms", "chunk_map"]


def step_size_terms(step: int, lr: float, beta1: float, beta2: float, correct_bias: bool) -> Tuple[float, float]:
    """(ss_base, ss_floor) as the reference forms them (:173-180).  Its step counter is an int32 TENSOR, so with bias
    correction `beta ** step`, the corrections and the quotient are fp32 0-dim tensors (and so is the product with the
    rms tensor or with the float 1e-3); without it the step size stays the python float lr."""
    if not correct_bias:
        return float(torch.tensor(lr, dtype=torch.float32)), lr * 1e-3
    t = torch.tensor(step, dtype=torch.int32)
    bias_correction1 = 1.0 - beta1 ** t
    bias_correction2 = 1.0 - beta2 ** t
    ss = lr * math.sqrt(bias_correction2) / bias_correction1        # fp32 tensor
    return float(ss), float(ss * 1e-3)


def chunk_map(numels, chunk: int):
    """(first_chunk per tensor, chunk -> tensor list): every tensor owns ceil(numel / chunk) consecutive chunks."""
    first, owner = [], []
    for i, n in enumerate(numels):
        first.append(len(owner))
        owner.extend([i] * ((n + chunk - 1) // chunk))
    return first, owner


class AdamWScale(Optimizer):

    def __init__(self, params: Iterable[nn.parameter.Parameter], lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-6, weight_decay: float = 0.0, kahan_sum: bool = False, foreach: bool = False,
                 correct_bias: bool = True, use_state_dtype: torch.dtype = None):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr} - should be >= 0.0")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter: {betas[0]} - should be in [0.0, 1.0)")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter: {betas[1]} - should be in [0.0, 1.0)")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps} - should be >= 0.0")
        assert not (foreach and use_state_dtype is not None), "foreach is not supported with use_state_dtype"
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, foreach=foreach, kahan_sum=kahan_sum,
                        correct_bias=correct_bias, use_state_dtype=use_state_dtype)
        super().__init__(params, defaults)
        self._chunk_cache = {}
        self._table_cache = {}

    @staticmethod
    def _rms(tensor):
        return tensor.norm(2) / (tensor.numel() ** 0.5)

    def _init_state(self, p, group):
        """State exactly as the reference creates it (:99-113)."""
        state = self.state[p]
        if "kahan_comp" not in state:
            state["step"] = torch.tensor(0, dtype=torch.int32, device=p.device)
            if group["use_state_dtype"] in [torch.float16, torch.bfloat16]:
                state["exp_avg"] = torch.zeros_like(p, device=p.device, dtype=group["use_state_dtype"])
                state["exp_avg_sq"] = torch.zeros_like(p, device=p.device, dtype=group["use_state_dtype"])
            else:
                state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if group["kahan_sum"] and p.dtype in [torch.float16, torch.bfloat16]:
                state["kahan_comp"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            else:
                state["kahan_comp"] = None
                group["kahan_sum"] = False
        if "step_host" not in state:                       # host copy of the counter (one sync after a checkpoint load)
            state["step_host"] = int(state["step"])
        return state

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            loss = closure()
        lib = _cabi.load()
        chunk = lib.b200t5_adamw_chunk_elems()
        for group in self.param_groups:
            buckets = {}
            steps = []
            active = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("AdamWScale does not support sparse gradients")
                _cabi.require_cuda(p, p.grad)
                state = self._init_state(p, group)
                state["step_host"] += 1
                steps.append(state["step"])
                if not (p.is_contiguous() and p.grad.is_contiguous() and state["exp_avg"].is_contiguous()
                        and state["exp_avg_sq"].is_contiguous()) or p.grad.dtype != p.dtype:
                    raise RuntimeError("AdamWScale needs contiguous parameters, gradients of the parameter dtype and contiguous states")
                active.append((p, state))
            # The reference decides the Kahan path by the GROUP flag, read after every state of the group exists: one fp32
            # parameter resets it for the whole group (:109-113), and then 16-bit parameters of that group -- which do own a
            # compensation tensor if they were initialised first -- take the plain update too (:127-152 pass group["kahan_sum"]).
            kahan_group = bool(group["kahan_sum"])
            for p, state in active:
                kahan = kahan_group and state["kahan_comp"] is not None
                key = (p.device, p.dtype, state["exp_avg"].dtype, kahan)
                buckets.setdefault(key, []).append((p, state))
            if steps:
                torch._foreach_add_(steps, 1)              # the state tensors keep the reference's meaning (:120)
            beta1, beta2 = group["betas"]
            for (device, p_dtype, s_dtype, kahan), items in buckets.items():
                self._launch(lib, chunk, device, p_dtype, s_dtype, kahan, items, group["lr"], beta1, beta2, group["eps"],
                             group["weight_decay"], group["correct_bias"])
        return loss

    def _launch(self, lib, chunk, device, p_dtype, s_dtype, kahan, items, lr, beta1, beta2, eps, weight_decay, correct_bias):
        numels = tuple(p.numel() for p, _ in items)
        cached = self._chunk_cache.get((device, numels))
        if cached is None:
            first, owner = chunk_map(numels, chunk)
            cached = (first, torch.tensor(owner, dtype=torch.int32).to(device), len(owner))
            self._chunk_cache[(device, numels)] = cached
        first, chunk_tensor, n_chunks = cached
        if n_chunks == 0:
            return
        terms = {}
        table = (_cabi.AdamwTensor * len(items))()
        neg_lr_wd = float(torch.tensor(-lr * weight_decay, dtype=torch.float32)) if weight_decay > 0.0 else 0.0
        for i, (p, state) in enumerate(items):
            t = state["step_host"]
            if t not in terms:
                terms[t] = step_size_terms(t, lr, beta1, beta2, correct_bias)
            d = table[i]
            d.p, d.g = p.data_ptr(), p.grad.data_ptr()
            d.m, d.v = state["exp_avg"].data_ptr(), state["exp_avg_sq"].data_ptr()
            d.comp = state["kahan_comp"].data_ptr() if kahan else None
            d.numel, d.first_chunk = numels[i], first[i]
            d.sqrt_numel = numels[i] ** 0.5
            d.ss_base, d.ss_floor = terms[t]
            d.neg_lr_wd = neg_lr_wd
        # descriptor table -> device through a cached PINNED staging buffer and a cached device buffer: a non-blocking copy
        # enqueued on the current stream.  (A pageable source made every bucket of every step wait for all prior work of the
        # stream -- ADVICE r1.)  The staging buffer is rewritten on the next step only after this copy has executed: an event
        # wait that has long completed by then.
        raw = bytes(table)
        slot = self._table_cache.get((device, numels, kahan))
        if slot is None or slot[0].numel() != len(raw):
            slot = [torch.empty(len(raw), dtype=torch.uint8, pin_memory=True), torch.empty(len(raw), dtype=torch.uint8, device=device), None]
            self._table_cache[(device, numels, kahan)] = slot
        pinned, dev_table, copied = slot
        if copied is not None:
            copied.synchronize()
        pinned.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
        dev_table.copy_(pinned, non_blocking=True)
        slot[2] = torch.cuda.Event()
        slot[2].record(torch.cuda.current_stream(device))
        nbytes = lib.b200t5_adamw_workspace_bytes(len(items), n_chunks)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        round_step_to_p = (not correct_bias) and p_dtype in (torch.float16, torch.bfloat16)
        rc = lib.b200t5_adamw_scale_step(dev_table.data_ptr(), len(items), chunk_tensor.data_ptr(), n_chunks, ws.data_ptr(),
                                         nbytes, _cabi.dtype_code(p_dtype), _cabi.dtype_code(s_dtype), 1 if kahan else 0,
                                         beta1, beta2, eps, 1 if round_step_to_p else 0, device.index,
                                         _cabi.stream_ptr(device))
        _cabi.check(rc, "b200t5_adamw_scale_step")
