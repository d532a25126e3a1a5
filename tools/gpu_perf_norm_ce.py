"""Developer tool (GPU box): HBM-roofline numbers for the RMSNorm and cross-entropy kernels.
    python tools/gpu_perf_norm_ce.py [--out gpurun_out/norm_ce_perf.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flasht5_b200  # noqa: E402,F401


def timeit(fn, flush, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def timeit_graph(make_call, nsets, reps=3):
    """Host-overhead-free timing for short kernels: capture one call per input set (the sets together exceed L2, so
    no call finds its input cached) in a CUDA graph, replay it, divide."""
    calls = [make_call(i) for i in range(nsets)]
    for c in calls:
        c()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for c in calls:
            c()
    torch.cuda.current_stream().wait_stream(st)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for c in calls:
            c()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / nsets)
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "norm_ce_perf.json"))
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:   # noqa: BLE001
        peak = 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = []
    # RMSNorm: the reference benchmark shape 16 x S x 768 bf16 (benchmarks/bench_layernorm.py) + FAT5 widths
    for rows, n in ((16 * 512, 768), (16 * 1024, 768), (32 * 1024, 512), (32 * 1024, 1024)):
        nsets = max(4, int(400e6 // (rows * n * 2 * 3)) + 1)
        xs = [torch.randn(rows, n, device=dev).to(torch.bfloat16) for _ in range(nsets)]
        w = torch.ones(n, device=dev, dtype=torch.bfloat16)
        dys = [torch.randn_like(xs[0]) for _ in range(nsets)]
        y, rstd = torch.ops.b200t5.rmsnorm_fwd(xs[0], w, 1e-6)
        t_f = timeit_graph(lambda i: (lambda: torch.ops.b200t5.rmsnorm_fwd(xs[i], w, 1e-6)), nsets)
        t_b = timeit_graph(lambda i: (lambda: torch.ops.b200t5.rmsnorm_bwd(dys[i], xs[i], w, rstd, 1e-6)), nsets)
        bf, bb = 2 * rows * n * 2, 3 * rows * n * 2
        res.append({"op": "rmsnorm", "rows": rows, "n": n, "fwd_ms": t_f, "bwd_ms": t_b, "fwd_gbs": bf / t_f / 1e6,
                    "bwd_gbs": bb / t_b / 1e6, "fwd_frac": bf / t_f / 1e6 / peak, "bwd_frac": bb / t_b / 1e6 / peak})
        print("rmsnorm %6d x %4d  fwd %.4f ms %6.0f GB/s (%.2f)   bwd %.4f ms %6.0f GB/s (%.2f)" %
              (rows, n, t_f, bf / t_f / 1e6, bf / t_f / 1e6 / peak, t_b, bb / t_b / 1e6, bb / t_b / 1e6 / peak))
    # cross-entropy: 16 x S x 32768 bf16, z-loss on (benchmarks/bench_cross_entropy.py)
    for rows, V in ((16 * 512, 32768), (16 * 1024, 32768), (32 * 1024, 32768)):
        logits = torch.randn(rows, V, device=dev).to(torch.bfloat16)
        labels = torch.randint(0, V, (rows,), device=dev)
        dl = torch.full((rows,), 1.0 / rows, device=dev)
        losses, zl, lse = torch.ops.b200t5.ce_fwd(logits, labels, None, 0.0, 1.0, 1e-4, -100)
        t_f = timeit(lambda: torch.ops.b200t5.ce_fwd(logits, labels, None, 0.0, 1.0, 1e-4, -100), flush, iters=10)
        t_b = timeit(lambda: torch.ops.b200t5.ce_bwd(dl, logits, lse, labels, 0.0, 1.0, 1e-4, -100), flush, iters=10)
        t_bi = timeit(lambda: torch.ops.b200t5.ce_bwd_inplace(dl, logits, lse, labels, 0.0, 1.0, 1e-4, -100), flush, iters=10)
        bf, bb = rows * V * 2, 2 * rows * V * 2
        res.append({"op": "cross_entropy", "rows": rows, "vocab": V, "fwd_ms": t_f, "bwd_ms": t_b, "bwd_inplace_ms": t_bi,
                    "fwd_gbs": bf / t_f / 1e6, "bwd_gbs": bb / t_b / 1e6, "fwd_frac": bf / t_f / 1e6 / peak,
                    "bwd_frac": bb / t_b / 1e6 / peak})
        print("ce      %6d x %5d fwd %.4f ms %6.0f GB/s (%.2f)   bwd %.4f ms %6.0f GB/s (%.2f)   inplace %.4f ms" %
              (rows, V, t_f, bf / t_f / 1e6, bf / t_f / 1e6 / peak, t_b, bb / t_b / 1e6, bb / t_b / 1e6 / peak, t_bi))
        del logits
        torch.cuda.empty_cache()
    # T5 bias producer: table (32, H) -> bias (1, H, S, S) bf16 and back
    from flasht5_b200.positional_encoding import RelativePositionalEncoding
    for H, S in ((8, 1024), (12, 1024), (16, 4096)):
        pe = RelativePositionalEncoding(32, 128, H, S, bidirectional=True).to(dev)
        bias = pe.compute_bias(S, S, dtype=torch.bfloat16)
        nbytes = H * S * S * 2
        nsets = max(3, int(300e6 // nbytes) + 1)
        gbs = [torch.randn_like(bias) for _ in range(nsets)]
        lut = pe._bucket_lut(-(S - 1), S - 1, dev)
        w = pe.relative_attention_bias.weight.detach()
        t_f = timeit_graph(lambda i: (lambda: torch.ops.b200t5.t5_bias_fwd(w, lut, S - 1, None, None, S, S, torch.bfloat16)), nsets)
        t_b = timeit_graph(lambda i: (lambda: torch.ops.b200t5.t5_bias_bwd(gbs[i], lut, S - 1, None, None, 32)), nsets)
        res.append({"op": "t5_bias", "H": H, "S": S, "fwd_ms": t_f, "bwd_ms": t_b, "fwd_gbs": nbytes / t_f / 1e6,
                    "bwd_gbs": nbytes / t_b / 1e6, "fwd_frac": nbytes / t_f / 1e6 / peak, "bwd_frac": nbytes / t_b / 1e6 / peak})
        print("t5bias  H=%2d S=%4d     fwd %.4f ms %6.0f GB/s (%.2f)   bwd %.4f ms %6.0f GB/s (%.2f)" %
              (H, S, t_f, nbytes / t_f / 1e6, nbytes / t_f / 1e6 / peak, t_b, nbytes / t_b / 1e6, nbytes / t_b / 1e6 / peak))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"hbm_peak_gbs": peak, "results": res}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
