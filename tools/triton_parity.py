"""GPU-box tool: the new sm_100a kernels beside the REFERENCE Triton kernels on the same B200.

For every case: O, dQ, dK, dV, dBias of
    new          flasht5_b200.flash_attention_v2_bias (this repo, libb200t5.so)
    triton_def   the reference kernel with the config it ships for compute capability (10, 0): 32x32, 1 stage, 4 warps
                 (/root/reference/src/model/ops/flash_attention_v2_bias.py:322-323, :511-512)
    triton_a100  the same kernel with its A100 tile table forced through the sanctioned override of get_fwd_config /
                 get_bwd_config (:292, :487)
    eager_lowp   the reference's eager path attn_ref(upcast=False) + autograd (src/utils/attn_ref.py)
against an fp64 evaluation of the same formula (max |delta| and relative Frobenius), new against triton directly, the
reference test's own rule err <= 2 * err_eager + 1e-5 (tests/fa2_triton/test_fa2_bias.py:28,64-67), and forward / backward
times of each (CUDA events, L2 flushed between iterations), plus SDPA(attn_mask=bias) and upstream flash_attn_func without
bias as context bars (benchmarks/bench_fa2_bias.py:34-41).

The reference files are imported from the git-ignored staging copy baseline/_ref/ (tools/stage_reference.sh); nothing under
flasht5_b200/ imports them.

    python tools/triton_parity.py [--cases headline,c2,...] [--out gpurun_out/triton_parity.json] [--no-timing]
"""
import argparse
import json
import math
import os
import sys
import time
from unittest import mock

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_ROOT = os.path.join(ROOT, "baseline", "_ref")
sys.path.insert(0, REF_ROOT)

import flasht5_b200  # noqa: E402,F401
from flasht5_b200 import flash_attention_v2_bias as new_attn  # noqa: E402

DEV = "cuda:0"

# name: (B, H, M, N, D, bias kind, causal, backward?)   bias kind: None | "1H" (T5 layout) | "BH" (per-batch) | "1H_rand"
CASES = {
    "headline": (32, 8, 1024, 1024, 64, "1H", False, True),
    "headline_causal": (32, 8, 1024, 1024, 64, "1H", True, True),
    "c2": (32, 8, 512, 512, 64, "1H", False, True),
    "c3_enc": (16, 12, 1024, 1024, 64, "1H", False, True),
    "c3_dec_causal": (16, 12, 1024, 1024, 64, "1H", True, True),
    "c3_cross_nobias": (16, 12, 1024, 1024, 64, None, False, True),
    "c4_fwd": (8, 16, 4096, 4096, 64, "1H", False, False),
    "reftest_d128": (2, 4, 512, 612, 128, "BH", False, True),
    "reftest_d128_causal": (2, 4, 512, 612, 128, "BH", True, True),
    "reftest_d64": (2, 4, 1024, 1045, 64, "BH", False, True),
    "reftest_d64_causal": (2, 4, 1024, 1045, 64, "BH", True, True),
    "reftest_d64_bias1H": (2, 4, 1024, 1045, 64, "1H_rand", True, True),
}


def t5_bias(H, M, N, bidirectional, gen, dtype):
    """(1, H, M, N) Toeplitz bias from a (32, H) table ~ N(0, 0.5^2): the model's production layout."""
    from flasht5_b200.positional_encoding import RelativePositionalEncoding
    pe = RelativePositionalEncoding(32, 128, H, max(M, N), bidirectional=bidirectional).to(DEV)
    with torch.no_grad():
        pe.relative_attention_bias.weight.copy_(0.5 * torch.randn(32, H, generator=gen, device=DEV))
        b = pe.compute_bias(M, N, device=DEV) if hasattr(pe, "compute_bias") else None
    return b.to(dtype).contiguous()


def make_inputs(case, dtype, seed=1234):
    B, H, M, N, D, kind, causal, _ = case
    g = torch.Generator(device=DEV).manual_seed(seed)
    mk = lambda s: torch.randn(B, s, H, D, generator=g, device=DEV).to(dtype).permute(0, 2, 1, 3)  # noqa: E731
    q, k, v, do = mk(M), mk(N), mk(N), mk(M)
    bias = None
    if kind == "1H":
        try:
            bias = t5_bias(H, M, N, not causal, g, dtype)
        except Exception:   # producer signature differs: fall back to an explicit Toeplitz gather
            table = 0.5 * torch.randn(2 * max(M, N), H, generator=g, device=DEV)
            rel = torch.arange(N, device=DEV)[None, :] - torch.arange(M, device=DEV)[:, None] + max(M, N)
            bias = table[rel.clamp(0, 2 * max(M, N) - 1)].permute(2, 0, 1)[None].to(dtype).contiguous()
    elif kind == "1H_rand":
        bias = torch.randn(1, H, M, N, generator=g, device=DEV).to(dtype)
    elif kind == "BH":
        bias = torch.randn(B, H, M, N, generator=g, device=DEV).to(dtype)
    return q, k, v, bias, do


def oracle_fp64(q, k, v, bias, do, causal, scale, need_bwd):
    """fp64 evaluation, one batch element at a time (bounded memory); explicit backward formulas."""
    B, H, M, D = q.shape
    N = k.shape[2]
    o = torch.empty(B, H, M, D, dtype=torch.float64, device=DEV)
    dq = torch.zeros(B, H, M, D, dtype=torch.float64, device=DEV) if need_bwd else None
    dk = torch.zeros(B, H, N, D, dtype=torch.float64, device=DEV) if need_bwd else None
    dv = torch.zeros(B, H, N, D, dtype=torch.float64, device=DEV) if need_bwd else None
    db = torch.zeros(bias.shape, dtype=torch.float64, device=DEV) if (need_bwd and bias is not None) else None
    ms = torch.arange(M, device=DEV)[:, None]
    ns = torch.arange(N, device=DEV)[None, :]
    mask = (ms + N - M >= ns) if causal else None
    hs = 4 if M * N >= 4096 * 4096 else H        # head chunk for the 4096^2 case
    for b in range(B):
        for h0 in range(0, H, hs):
            sl = slice(h0, min(H, h0 + hs))
            qb, kb, vb = q[b, sl].double(), k[b, sl].double(), v[b, sl].double()
            s = torch.matmul(qb, kb.transpose(1, 2)) * scale
            if bias is not None:
                s = s + bias[b if bias.shape[0] > 1 else 0, sl if bias.shape[1] > 1 else slice(0, 1)].double()
            if mask is not None:
                s = s.masked_fill(~mask, float("-inf"))
            p = torch.softmax(s, dim=-1)
            p = torch.nan_to_num(p, nan=0.0)          # rows with no visible key
            ob = torch.matmul(p, vb)
            o[b, sl] = ob
            if need_bwd:
                dob = do[b, sl].double()
                dv[b, sl] = torch.matmul(p.transpose(1, 2), dob)
                dp = torch.matmul(dob, vb.transpose(1, 2))
                delta = (ob * dob).sum(-1, keepdim=True)
                ds = p * (dp - delta)
                dq[b, sl] = torch.matmul(ds, kb) * scale
                dk[b, sl] = torch.matmul(ds.transpose(1, 2), qb) * scale
                if db is not None:
                    bi = b if bias.shape[0] > 1 else 0
                    if bias.shape[1] > 1:
                        db[bi, sl] += ds
                    else:
                        db[bi, 0] += ds.sum(0)
    return o, dq, dk, dv, db


def err(x, ref):
    if x is None or ref is None:
        return None
    d = x.double() - ref
    return {"max_abs": float(d.abs().max()), "rel_f": float(d.norm() / (ref.norm() + 1e-300))}


def time_fn(fn, iters, warmup, flush):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default=",".join(CASES))
    ap.add_argument("--dtypes", default="bf16,fp16")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "triton_parity.json"))
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-timing", action="store_true")
    ap.add_argument("--no-triton", action="store_true")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)

    ref = None
    if not args.no_triton:
        from src.model.ops import flash_attention_v2_bias as ref      # the staged reference (Triton)
        from src.utils.attn_ref import attn_ref
    torch.backends.cuda.matmul.allow_tf32 = False
    flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    results = {"device": torch.cuda.get_device_name(0), "torch": torch.__version__, "cases": []}
    try:
        import triton
        results["triton"] = triton.__version__
    except Exception:
        pass
    t_start = time.time()

    def ref_cfg(kind, which, B, H, M, N, D, causal):
        fn = ref.get_fwd_config if kind == "fwd" else ref.get_bwd_config
        if which == "a100":
            with mock.patch.object(torch.cuda, "get_device_capability", lambda *a, **k: (8, 0)):
                return fn(B, H, M, N, D, causal)
        return fn(B, H, M, N, D, causal)

    for name in args.cases.split(","):
        case = CASES[name]
        B, H, M, N, D, kind, causal, need_bwd = case
        scale = 1.0
        for dn in args.dtypes.split(","):
            dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}[dn]
            if name in ("c4_fwd", "headline_causal", "c3_enc", "c3_dec_causal", "c3_cross_nobias") and dn == "fp16":
                continue                                      # the big shapes once (bf16): bounded run time
            rec = {"case": name, "shape": [B, H, M, N, D], "bias": kind, "causal": causal, "dtype": dn, "sm_scale": scale}
            q, k, v, bias, do = make_inputs(case, dtype)
            F = 4.0 * B * H * M * N * D / (2 if causal else 1)
            try:
                o64, dq64, dk64, dv64, db64 = oracle_fp64(q, k, v, bias, do, causal, scale, need_bwd)
                refs = {"o": o64, "dq": dq64, "dk": dk64, "dv": dv64, "dbias": db64}
                outs = {}

                def run_autograd(fn):
                    qq, kk, vv = (t.detach().requires_grad_(need_bwd) for t in (q, k, v))
                    bb = bias.detach().requires_grad_(need_bwd) if bias is not None else None
                    o = fn(qq, kk, vv, bb)
                    r = {"o": o.detach()}
                    if need_bwd:
                        ins = (qq, kk, vv) + ((bb,) if bb is not None else ())
                        g = torch.autograd.grad(o, ins, do)
                        r.update(dq=g[0], dk=g[1], dv=g[2], dbias=g[3] if bb is not None else None)
                    return r

                outs["new"] = run_autograd(lambda a, b_, c, d: new_attn(a, b_, c, d, causal, scale))
                if ref is not None:
                    for which in ("def", "a100"):
                        fcfg = ref_cfg("fwd", which, B, H, M, N, D, causal)
                        bcfg = ref_cfg("bwd", which, B, H, M, N, D, causal)
                        rec["triton_%s_cfg" % which] = {"fwd": list(fcfg), "bwd": list(bcfg)}
                        try:
                            with mock.patch.object(ref, "get_fwd_config", lambda *a, _c=fcfg: _c), \
                                    mock.patch.object(ref, "get_bwd_config", lambda *a, _c=bcfg: _c):
                                outs["triton_" + which] = run_autograd(
                                    lambda a, b_, c, d: ref.flash_attention_v2_bias(a, b_, c, d, causal, scale))
                        except Exception as e:      # e.g. out of shared memory for a tile config
                            rec["triton_%s_error" % which] = repr(e)[:300]
                    if M * N <= 1024 * 1100:
                        outs["eager_lowp"] = run_autograd(lambda a, b_, c, d: attn_ref(a, b_, c, d, sm_scale=scale, causal=causal, upcast=False))
                rec["err_vs_fp64"] = {impl: {t: err(r.get(t), refs[t]) for t in refs if r.get(t) is not None and refs[t] is not None}
                                      for impl, r in outs.items()}
                for which in ("def", "a100"):
                    if "triton_" + which in outs:
                        rec["new_vs_triton_" + which] = {
                            t: err(outs["new"][t], outs["triton_" + which][t].double())
                            for t in refs if outs["new"].get(t) is not None and outs["triton_" + which].get(t) is not None}
                if "eager_lowp" in outs:
                    rule = {}
                    for impl in outs:
                        if impl == "eager_lowp":
                            continue
                        rule[impl] = {t: rec["err_vs_fp64"][impl][t]["max_abs"] <= 2 * rec["err_vs_fp64"]["eager_lowp"][t]["max_abs"] + 1e-5
                                      for t in rec["err_vs_fp64"][impl] if t in rec["err_vs_fp64"]["eager_lowp"]}
                    rec["reference_test_rule_2x_eager"] = rule
                del refs, o64, dq64, dk64, dv64, db64, outs
                torch.cuda.empty_cache()

                if not args.no_timing:
                    timing = {}

                    def time_impl(label, fwd_fn, bwd_fn):
                        t = {}
                        try:
                            t["fwd_ms"] = time_fn(fwd_fn, args.iters, 3, flush)
                            t["fwd_tflops"] = F / t["fwd_ms"] / 1e9
                            if need_bwd and bwd_fn is not None:
                                t["bwd_ms"] = time_fn(bwd_fn, args.iters, 3, flush)
                                t["bwd_tflops"] = 2.5 * F / t["bwd_ms"] / 1e9
                                t["fwdbwd_tflops"] = 3.5 * F / (t["fwd_ms"] + t["bwd_ms"]) / 1e9
                        except Exception as e:
                            t["error"] = repr(e)[:300]
                        timing[label] = t

                    o_new, L_new = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, causal, scale)
                    time_impl("new", lambda: torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, causal, scale),
                              lambda: torch.ops.b200t5.attn_bias_bwd(o_new, do, q, k, v, bias, L_new, causal, scale))
                    if ref is not None:
                        for which in ("def", "a100"):
                            fcfg = ref_cfg("fwd", which, B, H, M, N, D, causal)
                            bcfg = ref_cfg("bwd", which, B, H, M, N, D, causal)
                            try:
                                o_t, L_t = torch.ops.flasht5.flash_attn_v2_fwd(q, k, v, bias, causal, scale, fcfg[0], fcfg[1], fcfg[3], fcfg[2])
                                time_impl("triton_" + which,
                                          lambda: torch.ops.flasht5.flash_attn_v2_fwd(q, k, v, bias, causal, scale, fcfg[0], fcfg[1], fcfg[3], fcfg[2]),
                                          lambda: torch.ops.flasht5.flash_attn_v2_bwd(o_t, do, q, k, v, bias, L_t, causal, scale, bcfg[0], bcfg[1], bcfg[3], bcfg[2]))
                            except Exception as e:
                                timing["triton_" + which] = {"error": repr(e)[:300]}
                        # context bars of the reference benchmark
                        import torch.nn.functional as Fn
                        qs, ks, vs = (t.detach().requires_grad_(need_bwd) for t in (q, k, v))
                        bs_ = bias.detach() if bias is not None else None
                        if not (causal and bias is not None):     # SDPA rejects attn_mask together with is_causal
                            def sdpa_f():
                                return Fn.scaled_dot_product_attention(qs, ks, vs, attn_mask=bs_, is_causal=causal, scale=scale)
                            o_s = sdpa_f()
                            time_impl("sdpa_attn_mask", lambda: sdpa_f(),
                                      (lambda: torch.autograd.grad(o_s, (qs, ks, vs), do, retain_graph=True)) if need_bwd else None)
                        try:
                            from flash_attn import flash_attn_func
                            qf, kf, vf = (t.detach().permute(0, 2, 1, 3).requires_grad_(need_bwd) for t in (q, k, v))
                            dof = do.permute(0, 2, 1, 3)
                            o_f = flash_attn_func(qf, kf, vf, softmax_scale=scale, causal=causal)
                            time_impl("flash_attn_func_nobias", lambda: flash_attn_func(qf, kf, vf, softmax_scale=scale, causal=causal),
                                      (lambda: torch.autograd.grad(o_f, (qf, kf, vf), dof, retain_graph=True)) if need_bwd else None)
                        except Exception as e:
                            timing["flash_attn_func_nobias"] = {"error": repr(e)[:300]}
                        if bias is not None:
                            time_impl("new_nobias", lambda: torch.ops.b200t5.attn_bias_fwd(q, k, v, None, causal, scale), None)
                    rec["timing"] = timing
            except Exception as e:
                rec["error"] = repr(e)[:500]
            rec["t_wall_s"] = round(time.time() - t_start, 1)
            results["cases"].append(rec)
            print(json.dumps(rec)[:1500], flush=True)
            with open(args.out, "w") as f:
                json.dump(results, f, indent=1)
            torch.cuda.empty_cache()
    print("wrote", args.out)


if __name__ == "__main__":
    main()
