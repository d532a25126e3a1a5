"""GPU-box tool: a handful of small attention / RMSNorm / cross-entropy calls, meant to run under
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitizer_cases.py
(SURVEY.md section 5).  Small shapes only: the sanitizer slows kernels down by 10-100x."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flasht5_b200  # noqa: E402,F401
from flasht5_b200 import flash_attention_v2_bias, flash_attention_v2_rpe, fast_rms_layernorm, cross_entropy_loss  # noqa: E402

DEV = "cuda:0"
CASES = [  # B, H, M, N, D, bias, causal, dtype
    (1, 2, 256, 256, 64, "1H", False, torch.bfloat16),
    (2, 1, 200, 328, 64, "BH", True, torch.bfloat16),
    (1, 2, 130, 130, 32, "11", False, torch.float16),
    (1, 1, 256, 384, 128, "1H", True, torch.bfloat16),
    (1, 2, 256, 256, 64, None, False, torch.bfloat16),
    (1, 2, 256, 256, 64, "rpe", True, torch.bfloat16),
    (1, 2, 130, 135, 16, "1H", False, torch.bfloat16),     # N % 8 != 0: pointer path
]
only = sys.argv[1:] and [int(x) for x in sys.argv[1].split(",")]
for i, (B, H, M, N, D, kind, causal, dt) in enumerate(CASES):
    if only and i not in only:
        continue
    g = torch.Generator(device=DEV).manual_seed(i)
    mk = lambda s: torch.randn(B, s, H, D, generator=g, device=DEV).to(dt).permute(0, 2, 1, 3).requires_grad_(True)  # noqa: E731
    q, k, v = mk(M), mk(N), mk(N)
    do = torch.randn(B, M, H, D, generator=g, device=DEV).to(dt).permute(0, 2, 1, 3)
    if kind == "rpe":
        w = (0.5 * torch.randn(H, 32, generator=g, device=DEV)).requires_grad_(True)
        o = flash_attention_v2_rpe(q, k, v, w, 128, causal=causal, sm_scale=1.0)
        torch.autograd.grad(o, (q, k, v, w), do)
    else:
        bias = None
        if kind:
            shape = {"BH": (B, H, M, N), "1H": (1, H, M, N), "11": (1, 1, M, N)}[kind]
            bias = torch.randn(shape, generator=g, device=DEV).to(dt).requires_grad_(True)
        o = flash_attention_v2_bias(q, k, v, bias, causal, 1.0)
        torch.autograd.grad(o, (q, k, v) + ((bias,) if bias is not None else ()), do)
    torch.cuda.synchronize()
    print("case", i, "ok", flush=True)
x = torch.randn(300, 768, device=DEV, dtype=torch.bfloat16, requires_grad=True)
w = torch.ones(768, device=DEV, dtype=torch.bfloat16, requires_grad=True)
fast_rms_layernorm(x, w, 1e-6).sum().backward()
logits = torch.randn(64, 32128, device=DEV, dtype=torch.bfloat16, requires_grad=True)
labels = torch.randint(0, 32128, (64,), device=DEV)
cross_entropy_loss(logits, labels, lse_square_scale=1e-4)[0].sum().backward()
torch.cuda.synchronize()
print("sanitizer cases done")
