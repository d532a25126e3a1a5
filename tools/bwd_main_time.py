"""GPU-box tool: CUDA-event time of the backward MAIN kernel alone (library profile hook, kernel id 2) at the headline shape, with and
without a dense bias.   [B200T5_LIB=flasht5_b200/libb200t5_<variant>.so] python tools/bwd_main_time.py"""
import sys, torch, os
sys.path.insert(0, ".")
import flasht5_b200
DEV = "cuda:0"
B, H, S, D = 32, 8, 1024, 64
g = torch.Generator(device=DEV).manual_seed(1)
mk = lambda: torch.randn(B, S, H, D, generator=g, device=DEV).to(torch.bfloat16).permute(0, 2, 1, 3)
q, k, v, do = mk(), mk(), mk(), mk()
from flasht5_b200 import _cabi
for mode in ("bias", "nobias"):
    bias = torch.randn(1, H, S, S, generator=g, device=DEV).to(torch.bfloat16) if mode == "bias" else None
    o, L = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, False, 1.0)
    for i in range(3): torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, bias, L, False, 1.0)
    torch.cuda.synchronize()
    _cabi.profile_enable(True); _cabi.profile_collect()
    for i in range(10): torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, bias, L, False, 1.0)
    torch.cuda.synchronize()
    pr = [ms for kid, ms in _cabi.profile_collect() if kid == 2]
    _cabi.profile_enable(False)
    print(os.path.basename(os.environ.get("B200T5_LIB", "default")), mode, "main bwd kernel us: %.1f" % (1e3 * sum(pr) / len(pr)), flush=True)
