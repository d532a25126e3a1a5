"""Two B200s, NCCL: the batch-sharded operator equals the unsharded one.

Each rank runs the CUDA kernels on its half of the batch; the only exchange is dBias, all-reduced in fp32 BEFORE its single
rounding (torch.ops.b200t5.attn_bias_bwd_f32dbias + data_parallel.allreduce_dbias_f32) on a side stream through a communicator
capped at 8 CTAs (data_parallel.comm_group).  Skipped on a box with fewer than two GPUs; run with
    gpurun --gpus 2 -- python -m pytest tests/test_data_parallel_nccl.py -m gpu -q
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import flasht5_b200  # noqa: F401
    from flasht5_b200.data_parallel import allreduce_dbias_f32, comm_group, shard_batch
    dev = torch.device("cuda", rank)
    B, H, S, D = 6, 4, 512, 64
    g = torch.Generator().manual_seed(5)                     # same global problem on every rank
    mk = lambda: torch.randn(B, S, H, D, generator=g).to(torch.bfloat16).to(dev).permute(0, 2, 1, 3)   # noqa: E731
    q, k, v, do = mk(), mk(), mk(), mk()
    bias = torch.randn(1, H, S, S, generator=g).to(torch.bfloat16).to(dev)
    ops = torch.ops.b200t5
    # unsharded reference on this rank's GPU (same kernels)
    o_f, L_f = ops.attn_bias_fwd(q, k, v, bias, False, 1.0)
    dq_f, dk_f, dv_f, db_f32 = ops.attn_bias_bwd_f32dbias(o_f, do, q, k, v, bias, L_f, False, 1.0)
    a, b = shard_batch(B, rank, world)
    qs, ks, vs, dos = (t[a:b] for t in (q, k, v, do))
    o, L = ops.attn_bias_fwd(qs, ks, vs, bias, False, 1.0)
    dq, dk, dv, db32 = ops.attn_bias_bwd_f32dbias(o, dos, qs, ks, vs, bias, L, False, 1.0)
    grp = comm_group()
    side = torch.cuda.Stream(device=dev)
    db = allreduce_dbias_f32(db32, torch.bfloat16, side, grp)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    relf = lambda x, y: float((x.double() - y.double()).norm() / y.double().norm())   # noqa: E731
    checks = {
        "comm_group": grp is not None,                       # torch 2.11 + NCCL 2.28: per-communicator CTA cap available
        "o": torch.equal(o, o_f[a:b]), "lse": torch.equal(L, L_f[a:b]),
        "dk": torch.equal(dk, dk_f[a:b]), "dv": torch.equal(dv, dv_f[a:b]),    # batch elements are independent, sums in fixed order
        "dq": relf(dq, dq_f[a:b]) < 3e-3,                    # 16-bit L2 reduce-add of the key blocks: the order varies from run to run
        "dbias": relf(db, db_f32) < 6e-3,                    # one bf16 rounding of the fp32 sum (2^-9) + the 16-bit group sums inside
    }
    with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
        f.write("1" if all(checks.values()) else "failed: %s (dq %g, dbias %g)" % ([k_ for k_, v_ in checks.items() if not v_], relf(dq, dq_f[a:b]), relf(db, db_f32)))
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_sharded_equals_unsharded_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").read_text() == "1", (tmp_path / "ok0").read_text()
    assert (tmp_path / "ok1").read_text() == "1", (tmp_path / "ok1").read_text()
