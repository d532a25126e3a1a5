"""Developer tool (GPU box): print the per-phase clock64 timeline of the forward kernel on one SM.
Needs the timing build of the library (attn_fwd.cu compiled with -DB200T5_FWD_TIMING):
    tools/build_variant.sh --headline hl_fwdtiming "-DB200T5_FWD_TIMING"
    B200T5_LIB=$PWD/flasht5_b200/libb200t5_hl_fwdtiming.so python tools/fwd_timeline.py [bias|nobias|rpe]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flasht5_b200  # noqa: E402,F401
from flasht5_b200 import flash_attention_rpe as rpe  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bias"
B, H, S, D = 32, 8, 1024, 64
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
mk = lambda: torch.randn(B, S, H, D, generator=g, device=dev).to(torch.bfloat16).permute(0, 2, 1, 3)   # noqa: E731
q, k, v = mk(), mk(), mk()
table = 0.5 * torch.randn(32, H, generator=g, device=dev)
lut, zero, lo, hi = rpe.bucket_lut(S, S, 32, 128, True, q.device)
bias = torch.ops.b200t5.t5_bias_fwd(table, lut, zero, None, None, S, S, torch.bfloat16)
band = torch.ops.b200t5.rpe_band(table, lut, zero, lo, hi, torch.bfloat16)
torch.cuda.synchronize()
print("== mode", mode, flush=True)
for rep in range(2):       # the second call is the warm one
    print("-- call", rep, flush=True)
    if mode == "bias":
        torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, False, 1.0)
    elif mode == "nobias":
        torch.ops.b200t5.attn_bias_fwd(q, k, v, None, False, 1.0)
    else:
        torch.ops.b200t5.attn_rpe_fwd(q, k, v, band, lo, hi, False, 1.0)
    torch.cuda.synchronize()
