"""Developer tool (GPU box): a few launches of the attention ops on one shape, for ncu captures.
    ncu ... python tools/prof_one.py --shape headline --reps 3 [--fwd-only]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import flasht5_b200  # noqa: E402,F401
from gpu_perf import SHAPES  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="headline")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--fwd-only", action="store_true")
a = ap.parse_args()
B, H, M, N, D, bk, causal, bwd = SHAPES[a.shape]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1234)
mk = lambda s: torch.randn(B, s, H, D, generator=g, device=dev).to(torch.bfloat16).permute(0, 2, 1, 3)  # noqa: E731
q, k, v, do = mk(M), mk(N), mk(N), mk(M)
bias = torch.randn(1, H, M, N, generator=g, device=dev).to(torch.bfloat16) if bk else None
for _ in range(a.reps):
    o, L = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, causal, 1.0)
    if bwd and not a.fwd_only:
        torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, bias, L, causal, 1.0)
torch.cuda.synchronize()
print("done")
