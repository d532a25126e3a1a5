"""Developer tool (GPU box): one short pass that (A) runs smoke(), (B) checks the in-kernel relative-position bias
mode against the golden vectors and against the dense-bias operator, (C) times fused vs dense at the headline and
long-sequence shapes.  Every result is appended to gpurun_out/oneshot.jsonl as soon as it exists, so a run that is
cut short still leaves what it measured.   usage: python tools/gpu_oneshot.py [--skip-timing]"""
import glob
import json
import os
import sys
import time
import traceback

T0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
OUT = open(os.path.join(ROOT, "gpurun_out", "oneshot.jsonl"), "a")


def log(**kw):
    kw["t"] = round(time.time() - T0, 1)
    OUT.write(json.dumps(kw) + "\n")
    OUT.flush()
    print(json.dumps(kw), flush=True)


import numpy as np   # noqa: E402
import torch         # noqa: E402

log(step="import", torch=torch.__version__, cuda=torch.cuda.is_available(),
    gpu=torch.cuda.get_device_name(0) if torch.cuda.is_available() else None)

from oracle import attn_bias_ref as orc                               # noqa: E402
import flasht5_b200                                                   # noqa: E402,F401
from flasht5_b200 import flash_attention_v2_rpe, flash_attention_v2_bias, _cabi   # noqa: E402
from flasht5_b200 import flash_attention_rpe as rpe                   # noqa: E402

DEV = "cuda:0"


def guarded(name, fn):
    try:
        fn()
        torch.cuda.synchronize()
        return True
    except Exception as e:   # noqa: BLE001
        log(step=name, ok=False, error=repr(e)[:500], tb=traceback.format_exc()[-800:])
        return False


# ---------------------------------------------------------------- A: smoke
def step_smoke():
    import __graft_entry__ as ge
    ge.smoke()
    log(step="smoke", ok=True)


guarded("smoke", step_smoke)


# ---------------------------------------------------------------- B: fused relative-position bias, correctness
def _t(a):
    return torch.from_numpy(np.asarray(a))


def run_rpe(z, causal, scale, maxd, fused):
    q, k, v, do = (_t(z[n]).to(torch.bfloat16).to(DEV) for n in ("q", "k", "v", "do"))
    w = _t(z["table"]).t().contiguous().to(DEV).requires_grad_(True)
    q.requires_grad_(True), k.requires_grad_(True), v.requires_grad_(True)
    o = flash_attention_v2_rpe(q, k, v, w, maxd, causal=causal, sm_scale=scale, fused=fused)
    dq, dk, dv, dw = torch.autograd.grad(o, (q, k, v, w), do)
    torch.cuda.synchronize()
    return {"o": o, "dq": dq, "dk": dk, "dv": dv, "dtable": dw.t()}


def step_golden():
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "rpe_*.npz"))):
        z = np.load(path)
        causal, scale, maxd = bool(z["causal"]), float(z["sm_scale"]), int(z["max_distance"])
        for fused in (False, True):
            got = run_rpe(z, causal, scale, maxd, fused)
            errs = {n: orc.error_metrics(got[n], _t(z[n]))[1] for n in ("o", "dq", "dk", "dv", "dtable")}
            log(step="golden", case=os.path.basename(path), fused=fused, relF=errs,
                ok=all(e < (4e-3 if n in ("o", "dv") else 1.2e-2) for n, e in errs.items()))


guarded("golden", step_golden)


def step_equal():
    shapes = [(2, 4, 512, 512, 64, False), (2, 4, 512, 512, 64, True), (1, 2, 300, 700, 128, False),
              (2, 2, 640, 384, 32, True), (1, 3, 130, 130, 16, False), (3, 8, 1024, 1024, 64, False),
              (1, 2, 2048, 2048, 64, True)]
    for (B, H, M, N, D, causal) in shapes:
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
        res = {}
        for i, name in enumerate(("o", "dq", "dk", "dv", "dw")):
            a, b = outs[False][i], outs[True][i]
            res[name] = {"equal": bool(torch.equal(a, b)), "relF": orc.error_metrics(b, a.double())[1]}
        log(step="equal", shape=[B, H, M, N, D, causal], res=res,
            ok=all(res[n]["equal"] for n in ("o", "dk", "dv")) and res["dq"]["relF"] < 4e-3 and res["dw"]["relF"] < 4e-3)


guarded("equal", step_equal)


# ---------------------------------------------------------------- C: timing, fused vs dense
def cuda_time(fn, warm=3, iters=10):
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


def step_timing():
    for (B, H, S, D, causal, do_bwd) in [(32, 8, 1024, 64, False, True), (32, 8, 1024, 64, True, True),
                                         (8, 16, 4096, 64, False, False)]:
        g = torch.Generator(device=DEV).manual_seed(1)
        mk = lambda: torch.randn(B, S, H, D, generator=g, device=DEV).to(torch.bfloat16).permute(0, 2, 1, 3)   # noqa: E731
        q, k, v, do = mk(), mk(), mk(), mk()
        table = (0.5 * torch.randn(32, H, generator=g, device=DEV))
        lut, zero, lo, hi = rpe.bucket_lut(S, S, 32, 128, not causal, q.device)
        bias = torch.ops.b200t5.t5_bias_fwd(table, lut, zero, None, None, S, S, torch.bfloat16)
        band = torch.ops.b200t5.rpe_band(table, lut, zero, lo, hi, torch.bfloat16)
        flops = 4.0 * B * H * S * S * D * (0.5 if causal else 1.0)
        t_dense = cuda_time(lambda: torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, causal, 1.0))
        t_fused = cuda_time(lambda: torch.ops.b200t5.attn_rpe_fwd(q, k, v, band, lo, hi, causal, 1.0))
        t_none = cuda_time(lambda: torch.ops.b200t5.attn_bias_fwd(q, k, v, None, causal, 1.0))
        rec = dict(step="timing", shape=[B, H, S, D, causal], fwd_ms=dict(dense=t_dense, fused=t_fused, nobias=t_none),
                   fwd_tflops=dict(dense=flops / t_dense / 1e9, fused=flops / t_fused / 1e9, nobias=flops / t_none / 1e9))
        if do_bwd:
            o, L = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, causal, 1.0)
            tb_dense = cuda_time(lambda: torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, bias, L, causal, 1.0))
            tb_fused = cuda_time(lambda: torch.ops.b200t5.attn_rpe_bwd(o, do, q, k, v, band, lut, zero, lo, hi, 32, L, causal, 1.0))
            tb_none = cuda_time(lambda: torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, None, L, causal, 1.0))
            rec["bwd_ms"] = dict(dense=tb_dense, fused=tb_fused, nobias=tb_none)
            rec["bwd_tflops"] = dict(dense=2.5 * flops / tb_dense / 1e9, fused=2.5 * flops / tb_fused / 1e9,
                                     nobias=2.5 * flops / tb_none / 1e9)
        log(**rec)
        del q, k, v, do, bias
        torch.cuda.empty_cache()


if "--skip-timing" not in sys.argv:
    guarded("timing", step_timing)
log(step="done", launches=_cabi.launch_count())
