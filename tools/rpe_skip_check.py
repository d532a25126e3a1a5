"""Developer tool (GPU box): the relative-position operator with and without the constant-tile dS skip
(B200T5_RPE_SKIP_CONST) against the reference goldens and against the composed dense route; then the backward timing."""
import glob
import json
import os
import sys
import time

T0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
OUT = open(os.path.join(ROOT, "gpurun_out", "rpe_skip_check.jsonl"), "a")


def log(**kw):
    kw["t"] = round(time.time() - T0, 1)
    OUT.write(json.dumps(kw) + "\n")
    OUT.flush()
    print(json.dumps(kw), flush=True)


import numpy as np   # noqa: E402
import torch         # noqa: E402
from oracle import attn_bias_ref as orc                               # noqa: E402
import flasht5_b200                                                   # noqa: E402,F401
from flasht5_b200 import flash_attention_v2_rpe                       # noqa: E402
from flasht5_b200 import flash_attention_rpe as rpe                   # noqa: E402

DEV = "cuda:0"
_t = lambda a: torch.from_numpy(np.asarray(a))   # noqa: E731
ok_all = True
for skip in ("0", "1", "2"):
    os.environ["B200T5_RPE_SKIP_CONST"] = skip
    try:
        for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "rpe_*.npz"))):
            z = np.load(path)
            causal, scale, maxd = bool(z["causal"]), float(z["sm_scale"]), int(z["max_distance"])
            q, k, v, do = (_t(z[n]).to(torch.bfloat16).to(DEV).requires_grad_(True) for n in ("q", "k", "v", "do"))
            w = _t(z["table"]).t().contiguous().to(DEV).requires_grad_(True)
            o = flash_attention_v2_rpe(q, k, v, w, maxd, causal=causal, sm_scale=scale, fused=True)
            dq, dk, dv, dw = torch.autograd.grad(o, (q, k, v, w), do.detach())
            torch.cuda.synchronize()
            errs = {n: orc.error_metrics(g, _t(z[n]))[1] for n, g in (("o", o), ("dq", dq), ("dk", dk), ("dv", dv), ("dtable", dw.t()))}
            ok = all(e < (4e-3 if n in ("o", "dv") else 1.2e-2) for n, e in errs.items())
            ok_all &= ok
            log(step="golden", skip=skip, case=os.path.basename(path), relF=errs, ok=ok)
        for (B, H, M, N, D, causal) in [(2, 4, 512, 512, 64, False), (1, 2, 1024, 1024, 64, True), (3, 8, 1024, 1024, 64, False),
                                        (1, 3, 700, 1300, 32, False)]:
            # (the causal M > N shape of tests/test_attention_rpe.py::test_cuda_constant_tile_skip is left to that test: every
            #  visible position falls into one bucket there, the exact table gradient is 0 and a relative error means nothing)
            g = torch.Generator().manual_seed(11)
            mk = lambda s: torch.randn(B, s, H, D, generator=g).to(torch.bfloat16).to(DEV).permute(0, 2, 1, 3)   # noqa: E731
            q, k, v, do = mk(M), mk(N), mk(N), mk(M)
            w = (0.5 * torch.randn(H, 32, generator=g)).to(DEV)
            outs = {}
            for fused in (False, True):
                qq, kk, vv, ww = (t.detach().clone().requires_grad_(True) for t in (q, k, v, w))
                o = flash_attention_v2_rpe(qq, kk, vv, ww, 128, causal=causal, sm_scale=1.0, fused=fused)
                outs[fused] = (o,) + torch.autograd.grad(o, (qq, kk, vv, ww), do)
            torch.cuda.synchronize()
            res = {n: {"equal": bool(torch.equal(outs[False][i], outs[True][i])),
                       "relF": orc.error_metrics(outs[True][i], outs[False][i].double())[1]}
                   for i, n in enumerate(("o", "dq", "dk", "dv", "dw"))}
            ok = all(res[n]["equal"] for n in ("o", "dk", "dv")) and res["dq"]["relF"] < 4e-3 and res["dw"]["relF"] < 4e-3
            ok_all &= ok
            log(step="equal", skip=skip, shape=[B, H, M, N, D, causal], res=res, ok=ok)
    except Exception as e:   # noqa: BLE001
        ok_all = False
        log(step="error", skip=skip, error=repr(e)[:400])
        break
log(step="summary", ok=ok_all)

if ok_all:
    B, H, S, D = 32, 8, 1024, 64
    g = torch.Generator(device=DEV).manual_seed(1)
    mk = lambda: torch.randn(B, S, H, D, generator=g, device=DEV).to(torch.bfloat16).permute(0, 2, 1, 3)   # noqa: E731
    q, k, v, do = mk(), mk(), mk(), mk()
    table = 0.5 * torch.randn(32, H, generator=g, device=DEV)
    lut, zero, lo, hi = rpe.bucket_lut(S, S, 32, 128, True, q.device)
    band = torch.ops.b200t5.rpe_band(table, lut, zero, lo, hi, torch.bfloat16)
    o, L = torch.ops.b200t5.attn_rpe_fwd(q, k, v, band, lo, hi, False, 1.0)
    for skip in ("0", "1", "2", "0", "1", "2"):
        os.environ["B200T5_RPE_SKIP_CONST"] = skip
        for _ in range(3):
            torch.ops.b200t5.attn_rpe_bwd(o, do, q, k, v, band, lut, zero, lo, hi, 32, L, False, 1.0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            torch.ops.b200t5.attn_rpe_bwd(o, do, q, k, v, band, lut, zero, lo, hi, 32, L, False, 1.0)
        b.record()
        torch.cuda.synchronize()
        log(step="timing", skip=skip, bwd_op_us=round(a.elapsed_time(b) / 20 * 1e3, 1))
log(step="done")
