"""GPU-box tool: forward + backward of a list of seeded cases against an fp64 evaluation on the GPU (max |delta| and
relative Frobenius per tensor), then fwd / bwd timings of the headline shape.  Used while bringing up a new kernel:
    python tools/bwd_check.py [--quick] [--no-timing]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import flasht5_b200  # noqa: E402,F401
from flasht5_b200 import flash_attention_rpe as rpe   # noqa: E402
from triton_parity import oracle_fp64, err            # noqa: E402

DEV = "cuda:0"
CASES = [  # B, H, M, N, D, bias kind, causal, dtype
    (1, 1, 128, 128, 64, None, False, torch.bfloat16), (1, 2, 256, 256, 64, "1H", False, torch.bfloat16),
    (2, 4, 512, 512, 64, "1H", False, torch.bfloat16), (2, 4, 512, 512, 64, "1H", True, torch.bfloat16),
    (3, 2, 300, 700, 64, "BH", False, torch.bfloat16), (2, 2, 640, 384, 32, "11", True, torch.bfloat16),
    (1, 3, 130, 130, 16, "1H", False, torch.float16), (2, 2, 1024, 1045, 64, "1H", True, torch.float16),
    (2, 2, 200, 328, 64, "B1", True, torch.bfloat16), (2, 3, 190, 70, 32, None, True, torch.float16),
    (2, 4, 512, 616, 128, "1H", False, torch.bfloat16), (1, 2, 384, 384, 128, None, True, torch.float16),
    (2, 4, 512, 512, 64, "rpe", True, torch.bfloat16), (2, 4, 512, 512, 64, "rpe", False, torch.bfloat16),
    (3, 5, 700, 300, 32, "rpe", False, torch.float16), (2, 8, 1024, 1024, 64, "rpe", False, torch.bfloat16),
    (9, 8, 1024, 1024, 64, "1H", False, torch.bfloat16),
]
if "--quick" in sys.argv:
    CASES = CASES[:4]
TOL = {torch.bfloat16: (4e-3, 1.2e-2), torch.float16: (6e-4, 2e-3)}      # (o, dv) , (dq, dk, dbias)
ok_all = True
for case in CASES:
    B, H, M, N, D, kind, causal, dt = case
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + M + N + D)
    mk = lambda s: torch.randn(B, s, H, D, generator=g, device=DEV).to(dt).permute(0, 2, 1, 3)   # noqa: E731
    q, k, v, do = mk(M), mk(N), mk(N), mk(M)
    rec = {"case": [str(x) for x in case]}
    try:
        if kind == "rpe":
            table = 0.5 * torch.randn(32, H, generator=g, device=DEV)
            lut, zero, lo, hi = rpe.bucket_lut(M, N, 32, 128, not causal, q.device)
            band = torch.ops.b200t5.rpe_band(table, lut, zero, lo, hi, dt)
            o, L = torch.ops.b200t5.attn_rpe_fwd(q, k, v, band, lo, hi, causal, 1.0)
            dq, dk, dv, dtab = torch.ops.b200t5.attn_rpe_bwd(o, do, q, k, v, band, lut, zero, lo, hi, 32, L, causal, 1.0)
            rel = torch.arange(N, device=DEV)[None, :] - torch.arange(M, device=DEV)[:, None]
            bidx = lut[(rel + zero).long()].long()                                   # (M, N) bucket of every position
            bias = table.to(dt)[bidx].permute(2, 0, 1)[None].contiguous()            # (1, H, M, N) in the io dtype
            r = oracle_fp64(q, k, v, bias, do, causal, 1.0, True)
            dtab_ref = torch.zeros(32, H, dtype=torch.float64, device=DEV)
            dtab_ref.index_add_(0, bidx.reshape(-1), r[4][0].permute(1, 2, 0).reshape(-1, H))
            outs = {"o": (o, r[0]), "dq": (dq, r[1]), "dk": (dk, r[2]), "dv": (dv, r[3]), "dtable": (dtab, dtab_ref)}
        else:
            bias = None
            if kind is not None:
                shape = {"BH": (B, H, M, N), "1H": (1, H, M, N), "11": (1, 1, M, N), "B1": (B, 1, M, N)}[kind]
                bias = torch.randn(shape, generator=g, device=DEV).to(dt)
            o, L = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, causal, 1.0)
            dq, dk, dv, db = torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, bias, L, causal, 1.0)
            r = oracle_fp64(q, k, v, bias, do, causal, 1.0, True)
            dbr = r[4]
            if bias is not None and kind == "B1":
                pass
            outs = {"o": (o, r[0]), "dq": (dq, r[1]), "dk": (dk, r[2]), "dv": (dv, r[3])}
            if bias is not None:
                outs["dbias"] = (db, dbr)
        torch.cuda.synchronize()
        good = True
        for name, (mine, ref) in outs.items():
            e = err(mine, ref)
            tol = TOL[dt][0] if name in ("o", "dv") else TOL[dt][1]
            fin = bool(torch.isfinite(mine.float()).all())
            rec[name] = [round(e["max_abs"], 5), float("%.3g" % e["rel_f"])]
            good &= fin and e["rel_f"] < tol
        rec["ok"] = good
    except Exception as ex:   # noqa: BLE001
        rec["error"] = repr(ex)[:300]
        rec["ok"] = False
        print(json.dumps(rec), flush=True)
        ok_all = False
        break                 # a CUDA error is sticky: stop here
    ok_all &= rec["ok"]
    print(json.dumps(rec), flush=True)
print("BWD_CHECK", "PASS" if ok_all else "FAIL", flush=True)

if "--no-timing" not in sys.argv and ok_all:
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

    B, H, S, D = 32, 8, 1024, 64
    g = torch.Generator(device=DEV).manual_seed(1)
    mk = lambda: torch.randn(B, S, H, D, generator=g, device=DEV).to(torch.bfloat16).permute(0, 2, 1, 3)   # noqa: E731
    q, k, v, do = mk(), mk(), mk(), mk()
    for kind, causal in (("1H", False), (None, False), ("1H", True), ("rpe", False)):
        if kind == "rpe":
            table = 0.5 * torch.randn(32, H, generator=g, device=DEV)
            lut, zero, lo, hi = rpe.bucket_lut(S, S, 32, 128, True, q.device)
            band = torch.ops.b200t5.rpe_band(table, lut, zero, lo, hi, torch.bfloat16)
            o, L = torch.ops.b200t5.attn_rpe_fwd(q, k, v, band, lo, hi, causal, 1.0)
            tf = cuda_time(lambda: torch.ops.b200t5.attn_rpe_fwd(q, k, v, band, lo, hi, causal, 1.0))
            tb = cuda_time(lambda: torch.ops.b200t5.attn_rpe_bwd(o, do, q, k, v, band, lut, zero, lo, hi, 32, L, causal, 1.0))
        else:
            bias = torch.randn(1, H, S, S, generator=g, device=DEV).to(torch.bfloat16) if kind else None
            o, L = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, causal, 1.0)
            tf = cuda_time(lambda: torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, causal, 1.0))
            tb = cuda_time(lambda: torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, bias, L, causal, 1.0))
        F = 68.72 / (2 if causal else 1)
        print(json.dumps({"timing": kind, "causal": causal, "fwd_us": round(tf, 1), "bwd_op_us": round(tb, 1), "fwd_tflops": round(F / tf * 1e3, 1),
                          "bwd_tflops": round(2.5 * F / tb * 1e3, 1), "fwdbwd_tflops": round(3.5 * F / (tf + tb) * 1e3, 1)}), flush=True)
