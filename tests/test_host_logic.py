"""CPU: host-side logic -- the operator surface mirrors the reference's, shape contracts of the fake
(meta) implementations, data-parallel sharding, bench.py arithmetic."""
import inspect
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def test_public_surface_mirrors_reference():
    import flasht5_b200 as ft
    sig = inspect.signature(ft.flash_attention_v2_bias)
    assert list(sig.parameters) == ["q", "k", "v", "bias", "causal", "sm_scale"]          # reference :274
    assert sig.parameters["causal"].default is False and sig.parameters["sm_scale"].default is None
    assert issubclass(ft.FlashAttentionAdditiveBias, torch.autograd.Function)
    assert list(inspect.signature(ft.fast_rms_layernorm).parameters) == ["X", "W", "eps"]  # rms_norm.py:285
    ce = inspect.signature(ft.cross_entropy_loss)
    assert list(ce.parameters) == ["logits", "labels", "precomputed_lse", "label_smoothing", "logit_scale",
                                   "lse_square_scale", "ignore_index", "inplace_backward", "process_group"]
    for op in ("attn_bias_fwd", "attn_bias_bwd", "rmsnorm_fwd", "rmsnorm_bwd", "ce_fwd", "ce_bwd", "ce_bwd_inplace"):
        assert hasattr(torch.ops.b200t5, op)


def test_fake_impls_shape_contract():
    import flasht5_b200  # noqa: F401
    B, H, M, N, D = 2, 3, 40, 56, 64
    q = torch.empty(B, M, H, D, device="meta", dtype=torch.bfloat16).permute(0, 2, 1, 3)
    k = torch.empty(B, N, H, D, device="meta", dtype=torch.bfloat16).permute(0, 2, 1, 3)
    bias = torch.empty(1, H, M, N, device="meta", dtype=torch.bfloat16)
    o, L = torch.ops.b200t5.attn_bias_fwd(q, k, k, bias, True, 1.0)
    assert o.shape == q.shape and o.stride() == q.stride() and o.dtype == q.dtype
    assert L.shape == (B, H, M) and L.dtype == torch.float32
    dq, dk, dv, ds = torch.ops.b200t5.attn_bias_bwd(o, o, q, k, k, bias, L, True, 1.0)
    assert dq.shape == q.shape and dk.shape == k.shape and dv.shape == k.shape and ds.shape == bias.shape
    dq, dk, dv, ds = torch.ops.b200t5.attn_bias_bwd(o, o, q, k, k, None, L, True, 1.0)
    assert ds.numel() == 0
    x = torch.empty(10, 512, device="meta", dtype=torch.bfloat16)
    w = torch.empty(512, device="meta", dtype=torch.float32)
    y, rstd = torch.ops.b200t5.rmsnorm_fwd(x, w, 1e-6)
    assert y.shape == x.shape and rstd.shape == (10,) and rstd.dtype == torch.float32
    dx, dw = torch.ops.b200t5.rmsnorm_bwd(x, x, w, rstd, 1e-6)
    assert dx.shape == x.shape and dw.shape == (512,) and dw.dtype == torch.float32
    lg = torch.empty(10, 1000, device="meta", dtype=torch.bfloat16)
    lb = torch.empty(10, device="meta", dtype=torch.long)
    losses, z, lse = torch.ops.b200t5.ce_fwd(lg, lb, None, 0.0, 1.0, 1e-4, -100)
    assert losses.shape == (10,) and z.shape == (10,) and lse.shape == (10,)
    assert torch.ops.b200t5.ce_bwd(losses, lg, lse, lb, 0.0, 1.0, 1e-4, -100).shape == lg.shape


def test_alignment_helper():
    from flasht5_b200.flash_attention_v2_bias import _aligned
    a = torch.zeros(2, 16, 4, 64, dtype=torch.bfloat16).permute(0, 2, 1, 3)
    assert _aligned(a)
    assert not _aligned(a[..., 1:])                   # base off by 2 bytes / non-multiple-of-8 stride
    assert not _aligned(torch.zeros(2, 4, 16, 64, dtype=torch.bfloat16).transpose(2, 3))
    assert _aligned(torch.zeros(1, 4, 16, 64, dtype=torch.bfloat16)[:, :, :, :])


def test_shard_batch_partitions_exactly():
    from flasht5_b200.data_parallel import shard_batch
    for B in (1, 7, 32, 256):
        for G in (1, 2, 3, 4, 8):
            spans = [shard_batch(B, r, G) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_batch(8, 2, 2)


def test_bench_flop_convention_and_reference_arm():
    sys.path.insert(0, ROOT)
    import bench
    # reference benchmark convention (benchmarks/bench_fa2_bias.py:10-13): fwd 4*B*S^2*H*D, bwd 2.5x
    assert bench.flops_fwd(32, 8, 1024, 1024, 64) == 4 * 32 * 8 * 1024 * 1024 * 64
    assert abs(bench.flops_fwd_bwd(32, 8, 1024, 1024, 64) / 1e9 - 240.5) < 0.1
    assert bench.flops_fwd(16, 12, 1024, 1024, 64, causal=True) * 2 == bench.flops_fwd(16, 12, 1024, 1024, 64)
    peak, src = bench.load_peaks()
    assert 1000 < peak < 2500 and src


@pytest.mark.timeout(600)
def test_bench_reference_arm_prints_contract_line():
    env = dict(os.environ)
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "1"], cwd=ROOT, env=env, text=True, timeout=580)
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "TFLOP/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["metric"].startswith("attention TFLOP/s fwd+bwd bf16")
