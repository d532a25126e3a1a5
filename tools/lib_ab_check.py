"""Developer tool (GPU box): run a fixed set of seeded attention cases (forward + backward) through whichever library
B200T5_LIB points at and either save the outputs (--save FILE) or compare them with a saved set (--compare FILE):
O, LSE-dependent dK, dV must be bit-identical between two builds that only differ in scheduling; dQ and dBias are
compared at 16-bit-rounding level.  With --tol (builds that change the arithmetic) every output
is compared at that level instead.  Also prints fwd / bwd times of the headline shape.
    python tools/lib_ab_check.py --save /tmp/base.pt ; B200T5_LIB=... python tools/lib_ab_check.py --compare /tmp/base.pt [--tol]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flasht5_b200  # noqa: E402,F401
from flasht5_b200 import flash_attention_rpe as rpe   # noqa: E402

DEV = "cuda:0"
CASES = [  # B, H, M, N, D, bias kind, causal, dtype
    (2, 4, 512, 512, 64, "1H", False, torch.bfloat16), (2, 4, 512, 512, 64, "1H", True, torch.bfloat16),
    (3, 2, 300, 700, 64, "BH", False, torch.bfloat16), (2, 2, 640, 384, 32, "11", True, torch.bfloat16),
    (1, 3, 130, 130, 16, "1H", False, torch.float16), (2, 2, 1024, 1045, 64, "1H", True, torch.float16),
    (2, 4, 512, 616, 128, "1H", False, torch.bfloat16), (1, 2, 384, 384, 128, None, True, torch.float16),
    (40, 8, 1024, 1024, 64, "1H", False, torch.bfloat16), (2, 4, 512, 512, 64, "rpe", True, torch.bfloat16),
    (3, 5, 700, 300, 128, "rpe", False, torch.float16),
]


def run(case):
    B, H, M, N, D, kind, causal, dt = case
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + M + N + D)
    mk = lambda s: torch.randn(B, s, H, D, generator=g, device=DEV).to(dt).permute(0, 2, 1, 3)   # noqa: E731
    q, k, v, do = mk(M), mk(N), mk(N), mk(M)
    if kind == "rpe":
        table = 0.5 * torch.randn(32, H, generator=g, device=DEV)
        lut, zero, lo, hi = rpe.bucket_lut(M, N, 32, 128, not causal, q.device)
        band = torch.ops.b200t5.rpe_band(table, lut, zero, lo, hi, dt)
        o, L = torch.ops.b200t5.attn_rpe_fwd(q, k, v, band, lo, hi, causal, 1.0)
        dq, dk, dv, db = torch.ops.b200t5.attn_rpe_bwd(o, do, q, k, v, band, lut, zero, lo, hi, 32, L, causal, 1.0)
    else:
        bias = None
        if kind is not None:
            shape = {"BH": (B, H, M, N), "1H": (1, H, M, N), "11": (1, 1, M, N)}[kind]
            bias = torch.randn(shape, generator=g, device=DEV).to(dt)
        o, L = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, causal, 1.0)
        dq, dk, dv, db = torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, bias, L, causal, 1.0)
    torch.cuda.synchronize()
    return [t.detach().cpu() for t in (o, L, dk, dv, dq, db)]


def cuda_time(fn, warm=3, iters=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


# headline-only variant libraries (tools/build_variant.sh --headline) carry the D = 64 / bf16 instantiations only
HEADLINE_ONLY = "hl_" in os.path.basename(os.environ.get("B200T5_LIB", ""))
keep = [i for i, c in enumerate(CASES) if not HEADLINE_ONLY or (c[4] == 64 and c[7] == torch.bfloat16)]
outs = {i: run(CASES[i]) for i in keep}
if sys.argv[1] == "--save":
    torch.save(outs, sys.argv[2])
    print("saved", len(outs), "cases")
else:
    ref = torch.load(sys.argv[2])
    ok = True
    for i in keep:
        c, a, b = CASES[i], ref[i], outs[i]
        exact = all(torch.equal(x, y) for x, y in zip(a[:4], b[:4]))
        fin = lambda t: torch.nan_to_num(t.double(), neginf=0.0, posinf=0.0)   # noqa: E731  (LSE of empty rows is -inf)
        relf = lambda x, y: float((fin(x) - fin(y)).norm() / (fin(x).norm() + 1e-30)) if x.numel() else 0.0   # noqa: E731
        rel = [relf(x, y) for x, y in zip(a[4:], b[4:])]
        good = exact and all(r < 4e-3 for r in rel)
        if "--tol" in sys.argv:
            rel = [relf(x, y) for x, y in zip(a, b)]
            good = all(r < 4e-3 for r in rel)
        ok &= good
        print(json.dumps({"case": [str(x) for x in c], "o_L_dk_dv_bit_identical": exact, "dq_dbias_relF": rel, "ok": good}))
    print("AB_CHECK", "PASS" if ok else "FAIL")

B, H, S, D = 32, 8, 1024, 64
g = torch.Generator(device=DEV).manual_seed(1)
mk = lambda: torch.randn(B, S, H, D, generator=g, device=DEV).to(torch.bfloat16).permute(0, 2, 1, 3)   # noqa: E731
q, k, v, do = mk(), mk(), mk(), mk()
for kind in ("1H", None):
    bias = torch.randn(1, H, S, S, generator=g, device=DEV).to(torch.bfloat16) if kind else None
    o, L = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, False, 1.0)
    tf = cuda_time(lambda: torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, False, 1.0))
    tb = cuda_time(lambda: torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, bias, L, False, 1.0))
    print(json.dumps({"lib": os.environ.get("B200T5_LIB", "default"), "bias": kind, "fwd_us": round(tf, 1), "bwd_op_us": round(tb, 1),
                      "fwd_tflops": round(68.72 / tf * 1e3, 1), "fwdbwd_tflops": round(240.5 / (tf + tb) * 1e3, 1)}))
