"""GPU-box tool: run the headline backward through a -DB200T5_BWD_TIMING library (tools/build_variant.sh) and let the launcher
print the cycle-stamped timeline of block 777.   B200T5_LIB=flasht5_b200/libb200t5_hl_bwdtiming.so python tools/bwd_timeline.py [bias|nobias|rpe]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flasht5_b200  # noqa: E402,F401
from flasht5_b200 import flash_attention_rpe as rpe   # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bias"
DEV = "cuda:0"
B, H, S, D = 32, 8, 1024, 64
g = torch.Generator(device=DEV).manual_seed(1)
mk = lambda: torch.randn(B, S, H, D, generator=g, device=DEV).to(torch.bfloat16).permute(0, 2, 1, 3)   # noqa: E731
q, k, v, do = mk(), mk(), mk(), mk()
print("== mode", mode, flush=True)
for it in range(2):
    print("-- call", it, flush=True)
    if mode == "rpe":
        table = 0.5 * torch.randn(32, H, generator=g, device=DEV)
        lut, zero, lo, hi = rpe.bucket_lut(S, S, 32, 128, True, q.device)
        band = torch.ops.b200t5.rpe_band(table, lut, zero, lo, hi, torch.bfloat16)
        o, L = torch.ops.b200t5.attn_rpe_fwd(q, k, v, band, lo, hi, False, 1.0)
        torch.ops.b200t5.attn_rpe_bwd(o, do, q, k, v, band, lut, zero, lo, hi, 32, L, False, 1.0)
    else:
        bias = torch.randn(1, H, S, S, generator=g, device=DEV).to(torch.bfloat16) if mode == "bias" else None
        o, L = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, False, 1.0)
        torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, bias, L, False, 1.0)
    torch.cuda.synchronize()
