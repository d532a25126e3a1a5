"""GPU-box tool: L attention layers sharing one (1, H, S, S) bias, forward + backward through
autograd -- the plain operator (autograd adds L rounded dBias tensors) against flash_attention_v2_bias_shared (one fp32
accumulator filled by the finalize kernels, one gradient to autograd).  Prints time per stack and the dBias error of both
against an fp64 sum (small shape) .   python tools/shared_bias_perf.py [L]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flasht5_b200 import SharedBiasGrad, flash_attention_v2_bias, flash_attention_v2_bias_shared  # noqa: E402

DEV = "cuda:0"
L = int(sys.argv[1]) if len(sys.argv) > 1 else 12


def stack(B, H, S, D, shared, qkv, dos, bias0):
    bias = bias0.detach().requires_grad_(True)
    acc = SharedBiasGrad()
    leaves, outs = [], []
    for i in range(L):
        q, k, v = (t.detach().requires_grad_(True) for t in qkv[i])
        o = flash_attention_v2_bias_shared(q, k, v, bias, acc, False, 1.0) if shared else flash_attention_v2_bias(q, k, v, bias, False, 1.0)
        leaves += [q, k, v]
        outs.append(o)
    grads = torch.autograd.grad(outs, leaves + [bias], dos)
    return grads[-1]


for (B, H, S, D) in [(32, 8, 1024, 64), (16, 12, 1024, 64)]:
    g = torch.Generator(device=DEV).manual_seed(3)
    mk = lambda: torch.randn(B, S, H, D, generator=g, device=DEV).to(torch.bfloat16).permute(0, 2, 1, 3)   # noqa: E731
    qkv = [(mk(), mk(), mk()) for _ in range(L)]
    dos = [mk() for _ in range(L)]
    bias0 = (0.5 * torch.randn(1, H, S, S, generator=g, device=DEV)).to(torch.bfloat16)
    res = {"shape": [B, H, S, S, D], "layers": L}
    for shared in (False, True):
        for _ in range(2):
            stack(B, H, S, D, shared, qkv, dos, bias0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            gb = stack(B, H, S, D, shared, qkv, dos, bias0)
        e1.record()
        torch.cuda.synchronize()
        res["shared_ms" if shared else "plain_ms"] = e0.elapsed_time(e1) / 5
        res["shared_grad" if shared else "plain_grad"] = gb
    # error of the two bias gradients against an fp32-accumulated sum of per-layer UNROUNDED gradients
    ref = torch.zeros(1, H, S, S, dtype=torch.float64, device=DEV)
    for i in range(L):
        q, k, v = qkv[i]
        o, Ls = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias0, False, 1.0)
        ref += torch.ops.b200t5.attn_bias_bwd_f32dbias(o, dos[i], q, k, v, bias0, Ls, False, 1.0)[3].double()
    for key in ("plain", "shared"):
        gb = res.pop(key + "_grad")
        res[key + "_relF_vs_fp32_sum"] = float((gb.double() - ref).norm() / ref.norm())
    print(json.dumps(res), flush=True)
