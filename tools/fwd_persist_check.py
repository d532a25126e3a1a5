"""Developer tool (GPU box): the persistent forward schedule (attn_fwd_persist.cu) and the two-query-tile forward
(attn_fwd_pingpong.cu) against the one-CTA-per-block kernel (attn_fwd.cu): outputs must be bit-identical (same arithmetic,
different scheduling); then time all three.
Results are appended to gpurun_out/persist_check.jsonl as they come.   usage: python tools/fwd_persist_check.py"""
import json
import os
import sys
import time

T0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
OUT = open(os.path.join(ROOT, "gpurun_out", "persist_check.jsonl"), "a")


def log(**kw):
    kw["t"] = round(time.time() - T0, 1)
    OUT.write(json.dumps(kw) + "\n")
    OUT.flush()
    print(json.dumps(kw), flush=True)


import torch   # noqa: E402
import flasht5_b200  # noqa: E402,F401
from flasht5_b200 import flash_attention_rpe as rpe   # noqa: E402

DEV = "cuda:0"


def fwd(mode, q, k, v, bias, causal, scale, band=None, lo=0, hi=0):
    """mode: False / "base" = attn_fwd.cu, True / "persist" = attn_fwd_persist.cu, "pingpong" = attn_fwd_pingpong.cu"""
    os.environ["B200T5_FWD_PERSIST"] = "1" if mode in (True, "persist") else "0"
    os.environ["B200T5_FWD_PINGPONG"] = "1" if mode == "pingpong" else "0"
    if band is not None:
        return torch.ops.b200t5.attn_rpe_fwd(q, k, v, band, lo, hi, causal, scale)
    return torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, causal, scale)


# (B, H, M, N, D, bias kind, causal, dtype)
CASES = [
    (2, 4, 512, 512, 64, "1H", False, torch.bfloat16), (2, 4, 512, 512, 64, "1H", True, torch.bfloat16),
    (3, 2, 300, 700, 64, "BH", False, torch.bfloat16), (2, 2, 640, 384, 32, "11", True, torch.bfloat16),   # M > N causal: empty blocks
    (1, 3, 130, 130, 16, "1H", False, torch.float16), (2, 2, 1024, 1045, 64, "1H", True, torch.float16),   # N % 8 != 0: pointer path
    (2, 4, 512, 616, 128, "1H", False, torch.bfloat16), (1, 2, 384, 384, 128, None, True, torch.float16),
    (5, 8, 1024, 1024, 64, None, False, torch.bfloat16), (40, 8, 1024, 1024, 64, "1H", False, torch.bfloat16),  # > 296 items: several per CTA
    (9, 16, 2048, 2048, 64, "1H", True, torch.bfloat16), (37, 3, 256, 1280, 32, "BH", False, torch.bfloat16),
    (2, 4, 512, 512, 64, "rpe", False, torch.bfloat16), (2, 4, 512, 512, 64, "rpe", True, torch.bfloat16),
    (40, 8, 1024, 1024, 64, "rpe", False, torch.bfloat16), (3, 5, 700, 300, 128, "rpe", False, torch.float16),
    (11, 6, 1500, 1500, 16, "rpe", True, torch.bfloat16),
]


def make(case):
    B, H, M, N, D, kind, causal, dt = case
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + M + N + D)
    mk = lambda s: torch.randn(B, s, H, D, generator=g, device=DEV).to(dt).permute(0, 2, 1, 3)   # noqa: E731
    q, k, v = mk(M), mk(N), mk(N)
    bias = band = None
    lo = hi = 0
    if kind == "rpe":
        table = 0.5 * torch.randn(32, H, generator=g, device=DEV)
        lut, zero, lo, hi = rpe.bucket_lut(M, N, 32, 128, not causal, q.device)
        band = torch.ops.b200t5.rpe_band(table, lut, zero, lo, hi, dt)
    elif kind is not None:
        shape = {"BH": (B, H, M, N), "1H": (1, H, M, N), "11": (1, 1, M, N)}[kind]
        bias = torch.randn(shape, generator=g, device=DEV).to(dt)
    return q, k, v, bias, band, lo, hi


all_ok = True
all_ok_pp = True
for case in CASES:
    try:
        q, k, v, bias, band, lo, hi = make(case)
        o0, L0 = fwd(False, q, k, v, bias, case[6], 1.0, band, lo, hi)
        o1, L1 = fwd(True, q, k, v, bias, case[6], 1.0, band, lo, hi)
        o2, L2 = fwd(True, q, k, v, bias, case[6], 1.0, band, lo, hi)      # and once more: run-to-run determinism
        o3, L3 = fwd("pingpong", q, k, v, bias, case[6], 1.0, band, lo, hi)
        torch.cuda.synchronize()
        ok = bool(torch.equal(o0, o1) and torch.equal(L0, L1) and torch.equal(o1, o2) and torch.equal(L1, L2))
        ok_pp = bool(torch.equal(o0, o3) and torch.equal(L0, L3))
        log(step="equal_pingpong", case=[str(c) for c in case], ok=ok_pp, maxdiff=float((o0.float() - o3.float()).abs().max()))
        all_ok_pp &= ok_pp
        fin = bool(torch.isfinite(o1.float()).all())
        all_ok &= ok and fin
        log(step="equal", case=[str(c) for c in case], ok=ok, finite=fin,
            maxdiff=float((o0.float() - o1.float()).abs().max()))
    except Exception as e:   # noqa: BLE001
        all_ok = False
        log(step="equal", case=[str(c) for c in case], ok=False, error=repr(e)[:400])
        break
log(step="equal_summary", ok=all_ok, ok_pingpong=all_ok_pp)


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
    return a.elapsed_time(b) / iters


if all_ok or "--force-timing" in sys.argv:
    for case in [(32, 8, 1024, 1024, 64, "1H", False, torch.bfloat16), (32, 8, 1024, 1024, 64, "1H", True, torch.bfloat16),
                 (32, 8, 1024, 1024, 64, None, False, torch.bfloat16), (32, 8, 1024, 1024, 64, "rpe", False, torch.bfloat16),
                 (32, 8, 512, 512, 64, "1H", False, torch.bfloat16), (16, 12, 1024, 1024, 64, "1H", False, torch.bfloat16),
                 (8, 16, 4096, 4096, 64, "1H", False, torch.bfloat16), (8, 16, 4096, 4096, 64, "rpe", False, torch.bfloat16),
                 (16, 8, 1024, 1024, 128, "1H", False, torch.bfloat16)]:
        try:
            q, k, v, bias, band, lo, hi = make(case)
            B, H, M, N, D, kind, causal, dt = case
            flops = 4.0 * B * H * M * N * D * (0.5 if causal else 1.0)
            res = {}
            for mode in ("base", "persist", "pingpong", "base", "persist", "pingpong"):
                t = cuda_time(lambda: fwd(mode, q, k, v, bias, causal, 1.0, band, lo, hi))
                res.setdefault(mode, []).append(round(t * 1e3, 1))
            log(step="timing", case=[str(c) for c in case], us=res,
                tflops={n: round(flops / (min(v) * 1e-6) / 1e12, 1) for n, v in res.items()})
            del q, k, v, bias
            torch.cuda.empty_cache()
        except Exception as e:   # noqa: BLE001
            log(step="timing", case=[str(c) for c in case], error=repr(e)[:400])
            break
log(step="done")
