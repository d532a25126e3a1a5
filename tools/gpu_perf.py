"""Developer tool (GPU box): time the attention ops with CUDA events (L2 flushed between iterations).

    python tools/gpu_perf.py [--shapes headline,c2,...] [--iters 10] [--out gpurun_out/perf.jsonl]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flasht5_b200  # noqa: E402,F401

SHAPES = {
    # name: (B, H, M, N, D, bias, causal, bwd)
    "headline": (32, 8, 1024, 1024, 64, "1H", False, True),
    "headline_causal": (32, 8, 1024, 1024, 64, "1H", True, True),
    "headline_nobias": (32, 8, 1024, 1024, 64, None, False, True),
    "c2": (32, 8, 512, 512, 64, "1H", False, True),
    "c3": (16, 12, 1024, 1024, 64, "1H", False, True),
    "c3_causal": (16, 12, 1024, 1024, 64, "1H", True, True),
    "c4": (8, 16, 4096, 4096, 64, "1H", False, False),
    "d128": (16, 8, 1024, 1024, 128, "1H", False, True),
    "refbench_causal": (16, 12, 1024, 1024, 64, "1H", True, True),
}


def time_fn(fn, iters, warmup, flush):
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
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="headline,headline_causal,headline_nobias,c2,c3,c4,d128")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "perf.jsonl"))
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}[a.dtype]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    f = open(a.out, "a")
    print("%-18s %10s %10s %10s %10s %10s %10s" % ("shape", "fwd ms", "fwd TF/s", "bwd ms", "bwd TF/s", "f+b ms", "f+b TF/s"))
    for name in a.shapes.split(","):
        B, H, M, N, D, bk, causal, bwd = SHAPES[name]
        g = torch.Generator(device=dev).manual_seed(1234)
        mk = lambda s: torch.randn(B, s, H, D, generator=g, device=dev, dtype=torch.float32).to(dtype).permute(0, 2, 1, 3)  # noqa: E731
        q, k, v, do = mk(M), mk(N), mk(N), mk(M)
        bias = torch.randn(1, H, M, N, generator=g, device=dev).to(dtype) if bk else None
        F = 4.0 * B * H * M * N * D / (2 if causal else 1)
        o, L = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, causal, 1.0)
        fwd_med, fwd_min = time_fn(lambda: torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, causal, 1.0), a.iters, a.warmup, flush)
        rec = {"shape": name, "dims": [B, H, M, N, D], "bias": bk, "causal": causal, "dtype": a.dtype,
               "fwd_ms": fwd_med, "fwd_ms_min": fwd_min, "fwd_tflops": F / fwd_med / 1e9}
        if bwd:
            bwd_med, bwd_min = time_fn(lambda: torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, bias, L, causal, 1.0), a.iters, a.warmup, flush)
            rec.update({"bwd_ms": bwd_med, "bwd_ms_min": bwd_min, "bwd_tflops": 2.5 * F / bwd_med / 1e9,
                        "fb_ms": fwd_med + bwd_med, "fb_tflops": 3.5 * F / (fwd_med + bwd_med) / 1e9})
            print("%-18s %10.4f %10.1f %10.4f %10.1f %10.4f %10.1f" % (name, fwd_med, rec["fwd_tflops"], bwd_med, rec["bwd_tflops"], rec["fb_ms"], rec["fb_tflops"]))
        else:
            print("%-18s %10.4f %10.1f" % (name, fwd_med, rec["fwd_tflops"]))
        f.write(json.dumps(rec) + "\n")
        f.flush()
        del q, k, v, do, bias, o, L
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
