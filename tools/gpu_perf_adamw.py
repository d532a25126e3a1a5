"""Developer tool (GPU box): the fused AdamWScale step against the HBM roof and against the reference's own code paths.
A FAT5-small-like parameter set (about 110 M parameters in 258 tensors) is stepped with
  * flasht5_b200.AdamWScale (three launches per dtype group), and
  * an inline restatement of the reference's foreach arithmetic with torch._foreach ops (the reference module itself is not
    on the box), as the "what it replaces" number.
    python tools/gpu_perf_adamw.py [--out gpurun_out/adamw_perf.json] [--dtype fp32|bf16] [--kahan]
Algorithmic bytes per element: fp32 32 (p twice, g, m, v read; p, m, v written); bf16 + Kahan 20; bf16 16."""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flasht5_b200 import AdamWScale   # noqa: E402


def fat5_small_shapes():
    d, ff, h, dk, vocab, layers = 512, 1024, 8, 64, 32768, 12
    shapes = [(vocab, d), (32, h), (32, h)]
    for _ in range(layers):                       # encoder block
        shapes += [(d, h * dk)] * 4 + [(d,), (d, 2 * ff), (ff, d), (d,)]
    for _ in range(layers):                       # decoder block
        shapes += [(d, h * dk)] * 8 + [(d,), (d,), (d, 2 * ff), (ff, d), (d,)]
    shapes += [(d,), (d,), (vocab, d)]
    return shapes


def foreach_step(params, grads, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-6, wd=0.0):
    """The reference's _foreach_adamwscaled arithmetic (adamw_scaled.py:213-281) without Kahan, incl. its .item() per tensor."""
    torch._foreach_mul_(m, b1)
    torch._foreach_add_(m, grads, alpha=1 - b1)
    torch._foreach_mul_(v, b2)
    torch._foreach_addcmul_(v, grads, grads, 1 - b2)
    torch._foreach_copy_(grads, v)
    torch._foreach_sqrt_(grads)
    torch._foreach_add_(grads, eps)
    step_size = [torch.tensor(lr, dtype=torch.float32, device=p.device) for p in params]
    torch._foreach_mul_(step_size, [torch.tensor(math.sqrt(1 - b2 ** step.item()) / (1 - b1 ** step.item()), dtype=torch.float32,
                                                 device=p.device) for p in params])
    rms = torch._foreach_norm(params)
    torch._foreach_div_(rms, [torch.tensor(math.sqrt(p.numel())) for p in params])
    torch._foreach_maximum_(rms, 1e-3)
    torch._foreach_mul_(step_size, rms)
    torch._foreach_div_(grads, step_size)
    torch._foreach_addcdiv_(params, m, grads, value=-1)
    if wd > 0:
        torch._foreach_add_(params, params, alpha=-wd * lr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "adamw_perf.json"))
    ap.add_argument("--dtype", default="fp32")
    ap.add_argument("--kahan", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = {"fp32": torch.float32, "bf16": torch.bfloat16}[a.dtype]
    shapes = fat5_small_shapes()
    n = sum(math.prod(s) for s in shapes)
    g = torch.Generator(device=dev).manual_seed(0)
    params = [torch.nn.Parameter((0.02 * torch.randn(s, generator=g, device=dev)).to(dt)) for s in shapes]
    for p in params:
        p.grad = torch.randn(p.shape, generator=g, device=dev).to(dt)
    opt = AdamWScale(params, lr=1e-3, weight_decay=0.01, kahan_sum=a.kahan)

    def timed(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters

    ms = timed(opt.step)
    bpe = (32 if dt == torch.float32 else (20 if a.kahan else 16))
    res = {"params": n, "tensors": len(shapes), "dtype": a.dtype, "kahan": a.kahan, "fused_ms": ms,
           "fused_gbs": n * bpe / ms / 1e6, "bytes_per_element": bpe}
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:   # noqa: BLE001
        peak = 6546.6
    res["hbm_peak_gbs"], res["fused_frac"] = peak, res["fused_gbs"] / peak
    if not a.kahan:
        ps = [p.detach().clone() for p in params]
        gs = [p.grad.clone() for p in params]
        m = [torch.zeros_like(p) for p in ps]
        v = [torch.zeros_like(p) for p in ps]
        step = torch.tensor(1, dtype=torch.int32, device=dev)
        res["foreach_ms"] = timed(lambda: foreach_step(ps, gs, m, v, step), iters=5, warm=2)
        res["speedup_vs_foreach"] = res["foreach_ms"] / ms
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "a") as f:
        f.write(json.dumps(res) + "\n")
    print(json.dumps(res))


if __name__ == "__main__":
    main()
