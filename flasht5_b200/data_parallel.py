"""Data-parallel host logic for the attention operator.

The operator is embarrassingly parallel over batch: rank r of G owns a contiguous slice of the
batch and runs the kernels on it with no communication.  The single exchange is in the backward
and only because `bias` is one shared (1, H, M, N) tensor: dBias = sum over ranks of the local
dBias.  That is a plain all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_batch", "allreduce_dbias", "allreduce_dbias_overlapped", "allreduce_dbias_f32", "allreduce_dtable", "comm_group"]


def shard_batch(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, stop) of the batch rows rank `rank` owns; the first `global_batch % world_size` ranks
    get one extra row (same rule as torch.tensor_split)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(global_batch, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_dbias(dbias: Optional[torch.Tensor], group=None) -> Optional[torch.Tensor]:
    """Sum a batch-broadcast dBias over the data-parallel group.  The sum is done in fp32 (each rank's
    dBias is already a rounded 16-bit tensor; summing G of them in 16 bits would add G roundings) and
    cast back once.  No-op without an initialised process group or with a single rank."""
    if dbias is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dbias
    acc = dbias.float()
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc.to(dbias.dtype)


def allreduce_dbias_overlapped(dbias: Optional[torch.Tensor], comm_stream: "torch.cuda.Stream", group=None):
    """Same exchange as `allreduce_dbias`, enqueued on `comm_stream` so that it overlaps whatever the caller
    launches next on the current stream (in training: the rest of the backward pass).  Returns the reduced tensor;
    it is valid once the consumer stream has waited on `comm_stream`
    (`torch.cuda.current_stream().wait_stream(comm_stream)`)."""
    if dbias is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dbias
    comm_stream.wait_stream(torch.cuda.current_stream(dbias.device))
    with torch.cuda.stream(comm_stream):
        acc = dbias.float()
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
        out = acc.to(dbias.dtype)
    dbias.record_stream(comm_stream)
    return out


def allreduce_dtable(dtable: Optional[torch.Tensor], group=None) -> Optional[torch.Tensor]:
    """The exchange of the in-kernel relative-position operator (flash_attention_v2_rpe): its only shared gradient is
    the (num_buckets, H) table gradient -- 1 KB instead of the 33.5 MB dBias of the dense operator at the headline
    shape -- summed over ranks in fp32.  (Inside a DDP-wrapped model this is simply the embedding weight's bucket.)"""
    return allreduce_dbias(dtable, group)


def comm_group(max_ctas: int = 16):
    """A process group for the dBias exchange whose NCCL kernels are capped at `max_ctas` CTAs.  The exchange runs on a side
    stream BESIDE attention kernels that fill every SM (2 048-CTA grids); NCCL's default of up to 32 CTAs per collective took
    SMs and ~100 MB of HBM traffic from them (round 1: the backward kernel slowed from 0.306 to 0.341 ms at 2-8 ranks).  The
    message is 33.5 MB per step: a handful of CTAs moves it well within one step.  Measured at N = 2 on B200
    (profiles/r2c_n2_exchange_ab.txt, TFLOP/s of the whole job): no exchange 890 | cap 4: 820 | 6: 834 | 8: 836 / 544 (two runs:
    on the edge of being communication-bound) | 12: 844 | 16: 846 | 32: 656 | NCCL's default group: 565.  Falls back to the
    default group when the backend is not NCCL or the option is unavailable."""
    if not dist.is_available() or not dist.is_initialized():
        return None
    try:
        if dist.get_backend() != "nccl":
            return None
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = int(max_ctas)
        opts.config.min_ctas = 1
        return dist.new_group(ranks=list(range(dist.get_world_size())), backend="nccl", pg_options=opts)
    except Exception:      # noqa: BLE001  (older torch / NCCL without per-communicator config)
        return None


def allreduce_dbias_f32(dbias_f32: Optional[torch.Tensor], out_dtype: torch.dtype, comm_stream: "Optional[torch.cuda.Stream]" = None,
                        group=None) -> Optional[torch.Tensor]:
    """Exchange for the UNROUNDED fp32 dBias of `torch.ops.b200t5.attn_bias_bwd_f32dbias`: all-reduce it in place across the
    data-parallel group and round ONCE to `out_dtype` -- no widening cast before and no second rounding after the
    collective.  With `comm_stream` the collective and the cast are enqueued there (overlapping what the caller launches
    next); the result is valid once the consumer has waited on that stream."""
    if dbias_f32 is None:
        return None
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if comm_stream is None or not dbias_f32.is_cuda:
        if multi:
            dist.all_reduce(dbias_f32, op=dist.ReduceOp.SUM, group=group)
        return dbias_f32.to(out_dtype)
    comm_stream.wait_stream(torch.cuda.current_stream(dbias_f32.device))
    with torch.cuda.stream(comm_stream):
        if multi:
            dist.all_reduce(dbias_f32, op=dist.ReduceOp.SUM, group=group)
        out = dbias_f32.to(out_dtype)
    dbias_f32.record_stream(comm_stream)
    return out
