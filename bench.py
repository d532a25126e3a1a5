#!/usr/bin/env python
"""Benchmark of the hot path: FlashAttention-2 with additive T5 bias, forward + backward (dQ, dK, dV,
dBias), bf16, S = 1024 -- the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one forward + one backward pass of the operator over one batch of synthetic inputs.
Per-GPU workload (weak scaling): B=32, H=8, M=N=1024, D=64, bias (1,H,M,N): the per-GPU shape of
BASELINE.json configs[4] (FAT5-small, S=1024), which is the S=1024 shape the metric is quoted on.
FLOP convention = the reference benchmark's (benchmarks/bench_fa2_bias.py:10-13): fwd 4*B*H*M*N*D,
bwd 2.5x, fwd+bwd 3.5x.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "attention TFLOP/s fwd+bwd bf16 S=1024"
UNIT = "TFLOP/s"
# per-GPU workload
WB, WH, WS, WD = 32, 8, 1024, 64
SM_SCALE = 1.0                      # the reference's shipped configs use attention_scale 1.0
WORKLOAD = ("FAT5-small encoder self-attention fwd+bwd (dQ,dK,dV,dBias), per-GPU B=%d H=%d S=%d d=%d, "
            "bias (1,H,S,S), non-causal, sm_scale=1.0 (BASELINE.json configs[4] per-GPU shape)" % (WB, WH, WS, WD))


def flops_fwd(B, H, M, N, D, causal=False):
    return 4.0 * B * H * M * N * D / (2.0 if causal else 1.0)


def flops_fwd_bwd(B, H, M, N, D, causal=False):
    return 3.5 * flops_fwd(B, H, M, N, D, causal)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            pk = json.load(f)
        return float(pk["bf16_tflops"]), "measured burst (MEASURED_PEAKS.json bf16_tflops)"
    except Exception:   # noqa: BLE001
        return 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("attn_bwd_kernel_dram_bytes_per_launch")
    except Exception:   # noqa: BLE001
        return None


class ClockSampler:
    """Samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:   # noqa: BLE001
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:   # noqa: BLE001
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "power_w_max": max(power), "samples": len(sm),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's eager attention-with-bias (src/utils/attn_ref.py,
# restated in oracle/attn_bias_ref.py:attn_eager_lowp) + torch autograd on the host CPU cores.
# This is the ONLY place bench.py executes anything under oracle/.
# ---------------------------------------------------------------------------------------------
def cpu_eager_attention(sample_b: int, iters: int, warmup: int):
    import torch
    from oracle import attn_bias_ref as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1234)
    mk = lambda s: torch.randn(sample_b, s, WH, WD, generator=g).to(torch.bfloat16).permute(0, 2, 1, 3)  # noqa: E731
    q, k, v, do = mk(WS), mk(WS), mk(WS), mk(WS)
    bias = torch.randn(1, WH, WS, WS, generator=g).to(torch.bfloat16)
    for t in (q, k, v, bias):
        t.requires_grad_(True)
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + iters):
        t0 = time.perf_counter()
        o = orc.attn_eager_lowp(q, k, v, bias, False, SM_SCALE)
        torch.autograd.grad(o, (q, k, v, bias), do)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if times and time.perf_counter() - t_begin > 12.0:      # bounded: ~10-30 s of CPU work
            break
    iters = len(times)
    sec = sum(times) / len(times)
    tf = flops_fwd_bwd(sample_b, WH, WS, WS, WD) / sec / 1e12
    sample = ("B=%d of the per-GPU B=%d (same H=%d S=%d d=%d, bias (1,H,S,S)), %d warm-up + %d timed fwd+bwd "
              "iterations of the reference's eager bf16 attention (attn_ref semantics) + torch autograd, %d threads"
              % (sample_b, WB, WH, WS, WD, warmup, iters, cores))
    return tf, sec, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    iters = max(1, min(args.steps, 80))
    warm = max(1, min(args.warmup, 2))
    tf, sec, cores, sample = cpu_eager_attention(sample_b=8, iters=iters, warmup=warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": tf, "unit": UNIT, "n_gpus": args.gpus, "steps": iters,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference arm = the reference's eager PyTorch attention-with-bias on "
                   "the host CPU (oracle port of src/utils/attn_ref.py; the reference is pure Python and cannot be "
                   "compiled or shipped); each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": tf, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def bind_to_gpu_numa_node(local_rank: int):
    """Best effort: run this process (and therefore first-touch its pinned host buffers) on the NUMA node the GPU hangs off.
    Round 1's end-to-end leg swung 3.2x between two boxes (3.6 vs 11.3 ms per step) -- host buffers on the far socket.
    Returns a short description for the JSON line."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip()
        bus = out.lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]                                            # sysfs uses a 4-digit PCI domain
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa node unknown (single-node host or virtualised PCI)"
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return "numa node %d has no allowed cpus" % node
        os.sched_setaffinity(0, allowed)
        return "bound to numa node %d (%d cpus)" % (node, len(allowed))
    except Exception as e:   # noqa: BLE001
        return "not bound (%s)" % repr(e)[:80]


def load_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured copy (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:   # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def sibling_ops(torch, dev):
    """RMSNorm (rows = 32 x 1024, n = 512: FAT5-small) and cross-entropy + z-loss (rows = 32 768 is 2.1 GB of logits; a
    8 192-row slice is timed) forward / backward: algorithmic bytes over CUDA-event time,
    rotating over buffers larger than L2, against the measured HBM copy bandwidth."""
    from flasht5_b200 import fast_rms_layernorm, cross_entropy_loss
    hbm, hbm_src = load_hbm_peak()
    out = {"hbm_peak_gbs": hbm, "peak_source": hbm_src}

    def timed(fn, n=20):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    rows, n = 32 * 1024, 512
    xs = [torch.randn(rows, n, device=dev, dtype=torch.bfloat16) for _ in range(6)]       # 6 x 33.5 MB > L2 with dy / y
    dys = [torch.randn(rows, n, device=dev, dtype=torch.bfloat16) for _ in range(6)]
    w = torch.ones(n, device=dev, dtype=torch.bfloat16)
    ms = timed(lambda i: torch.ops.b200t5.rmsnorm_fwd(xs[i % 6], w, 1e-6))
    out["rmsnorm_fwd"] = {"rows": rows, "n": n, "ms": ms, "gbs": 2 * rows * n * 2 / ms / 1e6, "frac": 2 * rows * n * 2 / ms / 1e6 / hbm}
    y0, rstd0 = torch.ops.b200t5.rmsnorm_fwd(xs[0], w, 1e-6)
    ms = timed(lambda i: torch.ops.b200t5.rmsnorm_bwd(dys[i % 6], xs[i % 6], w, rstd0, 1e-6))
    out["rmsnorm_bwd"] = {"rows": rows, "n": n, "ms": ms, "gbs": 3 * rows * n * 2 / ms / 1e6, "frac": 3 * rows * n * 2 / ms / 1e6 / hbm}
    rows, V = 8192, 32768
    lg = [torch.randn(rows, V, device=dev, dtype=torch.bfloat16) for _ in range(2)]          # 2 x 537 MB
    labels = torch.randint(0, V, (rows,), device=dev)
    ms = timed(lambda i: torch.ops.b200t5.ce_fwd(lg[i % 2], labels, None, 0.0, 1.0, 1e-4, -100), n=10)
    out["ce_fwd"] = {"rows": rows, "vocab": V, "ms": ms, "gbs": rows * V * 2 / ms / 1e6, "frac": rows * V * 2 / ms / 1e6 / hbm}
    losses, z, lse = torch.ops.b200t5.ce_fwd(lg[0], labels, None, 0.0, 1.0, 1e-4, -100)
    dl = torch.ones(rows, device=dev)
    ms = timed(lambda i: torch.ops.b200t5.ce_bwd(dl, lg[i % 2], lse, labels, 0.0, 1.0, 1e-4, -100), n=10)
    out["ce_bwd"] = {"rows": rows, "vocab": V, "ms": ms, "gbs": 2 * rows * V * 2 / ms / 1e6, "frac": 2 * rows * V * 2 / ms / 1e6 / hbm}
    return out


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from flasht5_b200 import _cabi, flash_attention_v2_bias
    from flasht5_b200.data_parallel import allreduce_dbias, allreduce_dbias_f32, allreduce_dbias_overlapped, comm_group

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs the torchrun launch described in the docstring" % args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a B200: there is no CPU fallback")
    numa_note = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()
    if lib.b200t5_device_supported(local_rank) != 1:
        raise SystemExit("device is not sm_100: " + _cabi.last_error())

    B, H, S, D = WB, WH, WS, WD
    dtype = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    # The model's layout: (B, S, H, D) memory viewed as (B, H, S, D)  (modeling_flash_t5.py:281-283).
    # Three rotating input sets (3 x 185 MB > 126 MB L2) so no step finds its inputs in L2.
    NSETS = 3

    def mk():
        return torch.randn(B, S, H, D, generator=g, device=dev, dtype=torch.float32).to(dtype).permute(0, 2, 1, 3)
    sets = []
    for _ in range(NSETS):
        q, k, v, do = mk(), mk(), mk(), mk()
        bias = (0.5 * torch.randn(1, H, S, S, generator=g, device=dev)).to(dtype)
        sets.append((q, k, v, bias, do))
    F_step = flops_fwd_bwd(B, H, S, S, D)

    f32_exchange = world > 1 and args.exchange == "f32"

    def kernels(i):
        """The attention operators of one step: forward, then backward (dQ, dK, dV, dBias).  Nothing here allocates through the
        C ABI or synchronises, so the sequence is captured once per input set into a CUDA graph and replayed."""
        q, k, v, bias, do = sets[i % NSETS]
        o, L = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, False, SM_SCALE)
        if f32_exchange:
            dq, dk, dv, ds = torch.ops.b200t5.attn_bias_bwd_f32dbias(o, do, q, k, v, bias, L, False, SM_SCALE)   # unrounded fp32 dBias
        else:
            dq, dk, dv, ds = torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, bias, L, False, SM_SCALE)
        return o, dq, dk, dv, ds

    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    dp_group = comm_group(max_ctas=args.nccl_ctas) if (world > 1 and args.nccl_ctas > 0) else None

    def barrier():
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(comm_stream)
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):                  # eager warm-up: loads the kernels, sizes the allocator pools
        kernels(i)
    barrier()

    # One CUDA graph per input set (forward + the three backward launches): replay removes the host launch path and the
    # bubbles between dependent launches (~20 us of a 540 us step when launched eagerly from Python).
    graphs, launches_per_step = None, None
    if not args.no_graph:
        graphs = []
        c0 = _cabi.launch_count()
        for sidx in range(NSETS):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                outs = kernels(sidx)
            graphs.append((gr, outs))
        launches_per_step = (_cabi.launch_count() - c0) / NSETS
        torch.cuda.synchronize()
    buf_free = [None] * NSETS                             # N > 1: the exchange of the step that last used this graph's buffers

    def step(i):
        sidx = i % NSETS
        if graphs is not None:
            if buf_free[sidx] is not None:
                torch.cuda.current_stream(dev).wait_event(buf_free[sidx])
            graphs[sidx][0].replay()
            o, dq, dk, dv, ds = graphs[sidx][1]
        else:
            o, dq, dk, dv, ds = kernels(i)
        if world > 1 and args.exchange == "f32":
            # the one exchange of the path: the UNROUNDED fp32 dBias is summed over ranks on a side stream (NCCL capped at a
            # few CTAs so that it does not take SMs from the attention grids it overlaps) and rounded once afterwards
            ds = allreduce_dbias_f32(ds, dtype, comm_stream, dp_group)
        elif world > 1 and args.exchange == "bf16":       # developer A/B: round first, widen + all-reduce + round again (round 1)
            ds = allreduce_dbias_overlapped(ds, comm_stream, dp_group)
        if world > 1 and graphs is not None and args.exchange != "none":
            buf_free[sidx] = torch.cuda.Event()
            buf_free[sidx].record(comm_stream)
        return o, dq, dk, dv, ds

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    if graphs is None:
        _cabi.profile_enable(True)                          # per-kernel CUDA-event pairs inside the library (eager launches only)
        _cabi.profile_collect()
    launches0 = _cabi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(i)
    if world > 1:
        torch.cuda.current_stream(dev).wait_stream(comm_stream)     # every all-reduce is inside the timed region
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    if graphs is None:
        launches = _cabi.launch_count() - launches0
        prof = _cabi.profile_collect()
        _cabi.profile_enable(False)
        prof_note = "CUDA-event pairs around every launch of the timed region"
    else:
        launches = launches_per_step * args.steps          # kernel nodes replayed: counted when the graphs were captured
        # A replayed graph cannot carry event pairs: the per-kernel durations of the roofline come from an eager pass of the
        # same steps right after the timed region (same inputs, same clocks; clock sampler still running).
        n_prof = min(args.steps, 40)
        _cabi.profile_enable(True)
        _cabi.profile_collect()
        for i in range(n_prof):
            kernels(i)
        torch.cuda.synchronize()
        prof = _cabi.profile_collect()
        _cabi.profile_enable(False)
        prof_note = "CUDA-event pairs around every launch of %d eager steps run right after the timed (graph-replayed) region" % n_prof
    clocks = sampler.stop() if sampler else None

    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * F_step / (ms_step * 1e-3) / 1e12

    # ---- roofline of the dominant kernel (the fused backward), live CUDA-event times ----
    bwd_ms = [ms for kid, ms in prof if kid == 2 and ms > 0]
    fwd_ms = [ms for kid, ms in prof if kid == 1 and ms > 0]
    peak, peak_src = load_peaks()
    roofline = None
    if bwd_ms:
        avg = sum(bwd_ms) / len(bwd_ms)
        ach = 2.5 * flops_fwd(B, H, S, S, D) / (avg * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "attn_bwd_kernel_v3 (fused S^T/dP^T/dV/dK/dQ + dS^T out)", "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach / peak, "traffic": load_traffic(), "peak_source": peak_src,
                    "avg_launch_ms": avg, "launches_timed": len(bwd_ms), "timing": prof_note,
                    "algorithmic_flops_per_launch": 2.5 * flops_fwd(B, H, S, S, D)}
        if fwd_ms:
            favg = sum(fwd_ms) / len(fwd_ms)
            fach = flops_fwd(B, H, S, S, D) / (favg * 1e-3) / 1e12
            roofline["fwd_kernel"] = {"achieved": fach, "frac": fach / peak, "avg_launch_ms": favg}

    # ---- sustained leg: the same step loop for >= 2 s, against the SUSTAINED cuBLAS figure, with its own clock record ----
    sustained = None
    if not args.no_sustained:
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak_sus = float(json.load(f)["bf16_tflops_sustained"])
            peak_sus_src = "measured sustained (MEASURED_PEAKS.json bf16_tflops_sustained)"
        except Exception:   # noqa: BLE001
            peak_sus, peak_sus_src = 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"
        n_sus = max(args.steps, int(2200.0 / max(ms_step, 1e-3)) + 1)
        sampler2 = ClockSampler(local_rank) if rank == 0 else None
        if sampler2:
            sampler2.start()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(n_sus):
            step(i)
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(comm_stream)
        s1.record()
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sus_ms = float(ts.item()) / n_sus
        sus_val = world * F_step / (sus_ms * 1e-3) / 1e12
        sustained = {"value": sus_val, "unit": UNIT, "steps": n_sus, "seconds": float(ts.item()) * 1e-3, "ms_per_step": sus_ms,
                     "peak_tflops_per_gpu": peak_sus, "peak_source": peak_sus_src, "frac_of_sustained_peak": sus_val / (world * peak_sus),
                     "clocks": sampler2.stop() if sampler2 else None}

    # ---- the two sibling ops of the path (RMSNorm, cross-entropy + z-loss) against the HBM roof: one line each ----
    siblings = None
    if rank == 0 and world == 1 and not args.no_siblings:
        try:
            siblings = sibling_ops(torch, dev)
        except Exception as e:   # noqa: BLE001  (informational leg)
            siblings = {"error": repr(e)[:200]}

    # ---- e2e: host buffers -> public autograd API -> host results, copies inside the timed region ----
    # Every step moves q,k,v,bias,dO host->device and o,dq,dk,dv,dbias device->host (pinned memory).  Copies run on
    # their own streams and are double-buffered, so step i+1's upload and step i-1's download overlap step i's
    # kernels; the number is PCIe-bound (about 150 MB each way per step).
    e2e = None
    if not args.no_e2e:
        hq, hk, hv, hb, hdo = (x.detach().cpu().contiguous().pin_memory() for x in
                               (sets[0][0].permute(0, 2, 1, 3), sets[0][1].permute(0, 2, 1, 3),
                                sets[0][2].permute(0, 2, 1, 3), sets[0][3], sets[0][4].permute(0, 2, 1, 3)))
        outs_h = [[torch.empty_like(hq).pin_memory() for _ in range(4)] + [torch.empty_like(hb).pin_memory()]
                  for _ in range(2)]
        h2d = sum(x.numel() * x.element_size() for x in (hq, hk, hv, hb, hdo))
        d2h = sum(x.numel() * x.element_size() for x in outs_h[0])
        dev_in = [[torch.empty_like(x, device=dev) for x in (hq, hk, hv, hb, hdo)] for _ in range(2)]
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_comp = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        main = torch.cuda.current_stream(dev)
        for e in ev_comp + ev_out:
            e.record(main)

        def e2e_step(i):
            bsel = i % 2
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_comp[bsel])            # the kernels that last read this input buffer are done
                for dst, src in zip(dev_in[bsel], (hq, hk, hv, hb, hdo)):
                    dst.copy_(src, non_blocking=True)
                ev_in[bsel].record(s_in)
            main.wait_event(ev_in[bsel])
            dq_, dk_, dv_, db_, ddo = dev_in[bsel]
            q = dq_.permute(0, 2, 1, 3).requires_grad_(True)
            k = dk_.permute(0, 2, 1, 3).requires_grad_(True)
            v = dv_.permute(0, 2, 1, 3).requires_grad_(True)
            bb = db_.requires_grad_(True)
            o = flash_attention_v2_bias(q, k, v, bb, False, SM_SCALE)
            gq, gk, gv, gb = torch.autograd.grad(o, (q, k, v, bb), ddo.permute(0, 2, 1, 3))
            if world > 1:
                gb = allreduce_dbias(gb)
            ev_comp[bsel].record(main)
            for t_ in dev_in[bsel]:
                t_.requires_grad_(False)
            results = (o.permute(0, 2, 1, 3), gq.permute(0, 2, 1, 3), gk.permute(0, 2, 1, 3), gv.permute(0, 2, 1, 3), gb)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_comp[bsel])
                s_out.wait_event(ev_out[bsel])            # (host buffer reuse is ordered on this stream anyway)
                for dst, src in zip(outs_h[bsel], results):
                    dst.copy_(src, non_blocking=True)
                    src.record_stream(s_out)
                ev_out[bsel].record(s_out)

        e2e_steps = max(3, min(args.steps, 30))
        for i in range(2):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for i in range(e2e_steps):
            e2e_step(i)
        main.wait_stream(s_out)
        e1.record(main)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        dev_ms = e0.elapsed_time(e1)
        tt = torch.tensor([max(wall, dev_ms)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item()) / e2e_steps
        e2e = {"value": world * F_step / (e2e_ms * 1e-3) / 1e12, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
               "h2d_gbs": h2d / (e2e_ms * 1e-3) / 1e9, "d2h_gbs": d2h / (e2e_ms * 1e-3) / 1e9, "host_placement": numa_note,
               "api": "flasht5_b200.flash_attention_v2_bias + torch.autograd.grad; pinned host q,k,v,bias,dO in and "
                      "o,dq,dk,dv,dbias out every step, copies double-buffered on side streams"}

    # ---- the same workload through the in-kernel relative-position bias operator:
    #      bias from the (32, H) T5 table inside the kernels, gradient straight to the table.  Informational; the
    #      headline `value` above stays on the dense-bias operator BASELINE.json names. ----
    rpe_path = None
    if rank == 0 and world == 1:
        try:
            from flasht5_b200 import flash_attention_rpe as rpe
            table = 0.5 * torch.randn(32, H, generator=g, device=dev)
            lut, zero, lo, hi = rpe.bucket_lut(S, S, 32, 128, True, dev)
            band = torch.ops.b200t5.rpe_band(table, lut, zero, lo, hi, dtype)

            def rpe_step(i):
                q, k, v, _, do = sets[i % NSETS]
                o, L = torch.ops.b200t5.attn_rpe_fwd(q, k, v, band, lo, hi, False, SM_SCALE)
                return torch.ops.b200t5.attn_rpe_bwd(o, do, q, k, v, band, lut, zero, lo, hi, 32, L, False, SM_SCALE)
            for i in range(3):
                rpe_step(i)
            torch.cuda.synchronize()
            rpe_graphs = None
            if not args.no_graph:                              # same launch mode as the headline loop
                rpe_graphs = []
                for sidx in range(NSETS):
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr):
                        keep = rpe_step(sidx)
                    rpe_graphs.append((gr, keep))
                for gr, _ in rpe_graphs:
                    gr.replay()
                torch.cuda.synchronize()
            n_rpe = max(3, min(args.steps, 50))
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            for i in range(n_rpe):
                if rpe_graphs is not None:
                    rpe_graphs[i % NSETS][0].replay()
                else:
                    rpe_step(i)
            r1.record()
            torch.cuda.synchronize()
            rms = r0.elapsed_time(r1) / n_rpe
            rpe_path = {"value": F_step / (rms * 1e-3) / 1e12, "unit": UNIT, "ms_per_step": rms, "steps": n_rpe,
                        "api": "b200t5::attn_rpe_fwd + attn_rpe_bwd (flash_attention_v2_rpe): T5 bias computed in the "
                               "kernels from a (32, H) table, 32 buckets, max distance 128, bidirectional; "
                               "gradient = (32, H) table gradient"}
        except Exception as e:   # noqa: BLE001  (informational leg: never takes the bench line down)
            rpe_path = {"error": repr(e)[:200]}

    # ---- CPU baseline beside it (rank 0, N=1 only; bounded sample) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        tf, sec, cores, sample = cpu_eager_attention(sample_b=8, iters=80, warmup=1)      # stops after ~25 s
        cpu = {"value": tf, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "sec_per_iter": sec}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": WB * world, "seq_len": WS, "parallelism": "dp%d" % world,
                       "l2": "inputs rotate over %d buffer sets of 185 MB each (> 126 MB L2)" % NSETS,
                       "launch": "eager" if args.no_graph else "one CUDA graph per input set (fwd + 3 backward launches), replayed",
                       "exchange": ("none" if world == 1 else
                                    "DEVELOPER A/B --exchange %s --nccl-ctas %d: not the product configuration" % (args.exchange, args.nccl_ctas)
                                    if (args.exchange != "f32" or args.nccl_ctas != 16) else
                                    "NCCL all-reduce of the unrounded fp32 dBias (33.5 MB) every step on a side stream through a "
                                    "communicator capped at 16 CTAs, one rounding afterwards (overlaps the next step's kernels; "
                                    "all of them complete inside the timed region)")},
            "tokens_per_s": world * B * S / (ms_step * 1e-3),
            "frac_of_peak": value / (world * peak), "peak_tflops_per_gpu": peak, "peak_source": peak_src,
            "gpu_launches": int(launches), "launches_per_step": launches / args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "sustained": sustained,
            "rpe_path": rpe_path, "sibling_ops": siblings,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-siblings", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying the captured CUDA graphs")
    ap.add_argument("--exchange", default="f32", choices=["f32", "bf16", "none"], help="developer A/B of the N > 1 dBias exchange (the product is f32)")
    ap.add_argument("--nccl-ctas", type=int, default=16, help="CTA cap of the exchange communicator (0: NCCL's default group)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
