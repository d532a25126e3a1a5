"""CPU, world_size 2, gloo: the N>1 path of the operator harness -- shard the batch over ranks, no
collective in the forward, one all-reduce (dBias) in the backward.  The attention maths on the CPU
ranks is the ORACLE (test infrastructure); what is under test is the host-side sharding/exchange logic."""
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
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from flasht5_b200.data_parallel import allreduce_dbias, allreduce_dtable, shard_batch
    from oracle import attn_bias_ref as orc
    torch.manual_seed(0)                                   # same global problem on every rank
    B, H, M, N, D = 5, 2, 24, 40, 16
    q, k, v, do = torch.randn(B, H, M, D), torch.randn(B, H, N, D), torch.randn(B, H, N, D), torch.randn(B, H, M, D)
    bias = torch.randn(1, H, M, N)
    full = orc.attn_fwd_bwd(q, k, v, bias, do, True, 0.5)
    a, b = shard_batch(B, rank, world)
    loc = orc.attn_fwd_bwd(q[a:b], k[a:b], v[a:b], bias, do[a:b], True, 0.5)
    ok = True
    for i in (0, 2, 3, 4):                                  # O, dQ, dK, dV: purely local
        ok &= torch.allclose(loc[i], full[i][a:b], atol=1e-12)
    db = allreduce_dbias(loc[5].to(torch.bfloat16))         # 16-bit local dBias, fp32 exchange, one rounding
    ok &= db.dtype == torch.bfloat16
    ok &= torch.allclose(db.double(), full[5], atol=2e-2, rtol=2e-2)
    # the fp32 exchange of the unrounded dBias (attn_bias_bwd_f32dbias): summed before the single rounding
    from flasht5_b200.data_parallel import allreduce_dbias_f32, comm_group
    db32 = allreduce_dbias_f32(loc[5].float(), torch.bfloat16)
    ok &= db32.dtype == torch.bfloat16
    ok &= torch.equal(db32, full[5].float().to(torch.bfloat16)) or torch.allclose(db32.double(), full[5], atol=2e-2, rtol=1e-2)
    ok &= comm_group(8) is None                             # gloo: no NCCL communicator options, default group is used
    # exact in fp64
    db64 = loc[5].clone()
    dist.all_reduce(db64)
    ok &= torch.allclose(db64, full[5], atol=1e-12)
    # the relative-position operator: bias from a (32, H) table, the one shared gradient is the table's
    table = torch.randn(32, H, dtype=torch.float64)
    full_r = orc.attn_rpe_fwd_bwd(q, k, v, table, do, True, 0.5)
    loc_r = orc.attn_rpe_fwd_bwd(q[a:b], k[a:b], v[a:b], table, do[a:b], True, 0.5)
    for i in (0, 2, 3, 4):
        ok &= torch.allclose(loc_r[i], full_r[i][a:b], atol=1e-12)
    dt = allreduce_dtable(loc_r[5].float())
    ok &= dt.dtype == torch.float32 and dt.shape == (32, H)
    ok &= torch.allclose(dt.double(), full_r[5], atol=1e-4, rtol=1e-5)
    with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_batch_sharding_with_dbias_allreduce_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").read_text() == "1" and (tmp_path / "ok1").read_text() == "1"


def test_allreduce_is_noop_without_process_group():
    sys.path.insert(0, ROOT)
    from flasht5_b200.data_parallel import allreduce_dbias
    t = torch.randn(3)
    assert allreduce_dbias(t) is t and allreduce_dbias(None) is None
    from flasht5_b200.data_parallel import allreduce_dbias_f32, comm_group
    assert allreduce_dbias_f32(None, torch.bfloat16) is None and comm_group() is None
    assert allreduce_dbias_f32(t, torch.bfloat16).dtype == torch.bfloat16
