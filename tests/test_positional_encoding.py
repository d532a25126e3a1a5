"""T5 relative-position bias producer (first half): CPU checks of the host logic and the
oracle against golden vectors generated from the REFERENCE module (oracle/make_golden.py:gen_t5_bias), GPU checks
of the CUDA gather / segmented-sum kernels.  Bars: the forward gather is exact (index work: bit-identical values);
the backward sum is compared at 1e-5 relative (fp32 summation order)."""
import glob
import os

import numpy as np
import pytest
import torch

import flasht5_b200  # noqa: F401
from flasht5_b200.positional_encoding import RelativePositionalEncoding
from conftest import GOLDEN
from oracle import attn_bias_ref as orc

FILES = sorted(glob.glob(os.path.join(GOLDEN, "t5bias_*.npz")))
IDS = [os.path.basename(p)[:-4] for p in FILES]


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_fixtures_present():
    assert len(FILES) >= 3


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_oracle_bias_matches_reference_golden(path):
    z = np.load(path)
    b = orc.t5_bias(_t(z["table"]), int(z["M"]), int(z["N"]), bidirectional=bool(z["bidirectional"]),
                    num_buckets=int(z["num_buckets"]), max_distance=int(z["max_distance"]))
    assert torch.equal(b, _t(z["bias"]))


def test_module_bucket_function_known_answers():
    z = np.load(os.path.join(GOLDEN, "t5_buckets.npz"))
    rel = _t(z["rel"])
    f = RelativePositionalEncoding._relative_position_bucket
    assert f(rel.clone(), True, 32, 128).tolist() == z["bidirectional"].tolist()
    assert f(rel.clone(), False, 32, 128).tolist() == z["unidirectional"].tolist()


def test_module_surface_mirrors_reference():
    pe = RelativePositionalEncoding(32, 128, 8, 1024, bidirectional=False)
    assert list(pe.state_dict().keys()) == ["relative_attention_bias.weight"]        # checkpoints load unchanged
    assert pe.relative_attention_bias.weight.shape == (32, 8)
    lut = pe._bucket_lut(-5, 7, "cpu")
    assert lut.dtype == torch.int32 and lut.numel() == 13
    assert lut.tolist() == pe._relative_position_bucket(torch.arange(-5, 8), False, 32, 128).tolist()
    q = torch.zeros(2, 16, 8, 64)
    with pytest.raises((RuntimeError, NotImplementedError)):                          # no CPU fallback
        pe(q, q, q)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_cuda_bias_and_table_grad_match_reference_golden(path):
    z = np.load(path)
    M, N, H = int(z["M"]), int(z["N"]), z["table"].shape[1]
    pe = RelativePositionalEncoding(int(z["num_buckets"]), int(z["max_distance"]), H, max(M, N),
                                    bidirectional=bool(z["bidirectional"])).to("cuda:0")
    with torch.no_grad():
        pe.relative_attention_bias.weight.copy_(_t(z["table"]))
    bias = pe.compute_bias(M, N)
    assert bias.shape == (1, H, M, N) and bias.dtype == torch.float32 and bias.is_contiguous()
    assert torch.equal(bias.detach().cpu(), _t(z["bias"]))                            # gather: exact
    (dtable,) = torch.autograd.grad(bias, pe.relative_attention_bias.weight, _t(z["dbias"]).to("cuda:0"))
    mx, rf = orc.error_metrics(dtable, _t(z["dtable"]))
    assert rf < 1e-5, (mx, rf)
    # direct 16-bit emission == cast of the fp32 bias
    for dt in (torch.bfloat16, torch.float16):
        assert torch.equal(pe.compute_bias(M, N, dtype=dt), bias.detach().to(dt))


@pytest.mark.gpu
def test_cuda_forward_signature_and_full_size():
    """forward(q, k, v) -> (q, k, v, bias) with q: (B, S, H, D) as the model calls it (modeling_flash_t5.py:258-259);
    headline size S = 1024, H = 8, both directions, against the oracle producer."""
    for bidir in (True, False):
        pe = RelativePositionalEncoding(32, 128, 8, 1024, bidirectional=bidir).to("cuda:0")
        q = torch.zeros(2, 1024, 8, 64, device="cuda:0", dtype=torch.bfloat16)
        k = torch.zeros(2, 1000, 8, 64, device="cuda:0", dtype=torch.bfloat16)
        q2, k2, v2, bias = pe(q, k, k)
        assert q2 is q and k2 is k and v2 is k
        assert bias.shape == (1, 8, 1024, 1000) and bias.dtype == torch.bfloat16 and bias.is_contiguous()
        want = orc.t5_bias(pe.relative_attention_bias.weight.detach().cpu(), 1024, 1000, bidirectional=bidir)
        assert torch.equal(bias.detach().cpu(), want.to(torch.bfloat16))
        # backward against index_add on the CPU
        g = torch.randn(1, 8, 1024, 1000, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
        (dt,) = torch.autograd.grad(bias, pe.relative_attention_bias.weight, g.to("cuda:0"))
        rel = torch.arange(1000)[None, :] - torch.arange(1024)[:, None]
        bk = orc.t5_relative_position_bucket(rel, bidir)
        ref = torch.zeros(32, 8, dtype=torch.float64)
        ref.index_add_(0, bk.reshape(-1), g[0].double().permute(1, 2, 0).reshape(-1, 8))
        mx, rf = orc.error_metrics(dt, ref)
        assert rf < 1e-5, (mx, rf)


@pytest.mark.gpu
def test_cuda_randomized_positions():
    """use_randomized_position_encoding (positional_encoding.py:78-87): same sampling, same buckets."""
    H, L, M, N = 4, 512, 100, 120
    pe = RelativePositionalEncoding(32, 128, H, L, bidirectional=True, randomized_position=True).to("cuda:0")
    torch.manual_seed(123)
    bias = pe.compute_bias(M, N)
    torch.manual_seed(123)                                                            # replay the reference's sampling
    ci, _ = torch.sort(torch.randperm(L)[:M]); ci[0] = 0
    mi, _ = torch.sort(torch.randperm(L)[:N]); mi[0] = 0
    bk = orc.t5_relative_position_bucket(mi[None, :] - ci[:, None], True)
    want = pe.relative_attention_bias.weight.detach().cpu()[bk].permute(2, 0, 1).unsqueeze(0)
    assert torch.equal(bias.detach().cpu(), want)


@pytest.mark.gpu
def test_bias_feeds_attention_end_to_end():
    """table -> bias (CUDA) -> attention fwd+bwd (CUDA) -> dBias -> dTable (CUDA), against the fp64 oracle chain."""
    from flasht5_b200 import flash_attention_v2_bias
    B, H, S, D = 2, 4, 256, 64
    g = torch.Generator().manual_seed(5)
    pe = RelativePositionalEncoding(32, 128, H, S, bidirectional=False).to("cuda:0")
    with torch.no_grad():
        pe.relative_attention_bias.weight.copy_(0.5 * torch.randn(32, H, generator=g))
    mk = lambda: torch.randn(B, S, H, D, generator=g).to(torch.bfloat16)   # noqa: E731
    q, k, v, do = mk(), mk(), mk(), mk()
    qd, kd, vd = (t.to("cuda:0") for t in (q, k, v))
    _, _, _, bias = pe(qd, kd, vd)
    o = flash_attention_v2_bias(qd.permute(0, 2, 1, 3), kd.permute(0, 2, 1, 3), vd.permute(0, 2, 1, 3), bias, True, 1.0)
    (dtable,) = torch.autograd.grad(o, pe.relative_attention_bias.weight, do.to("cuda:0").permute(0, 2, 1, 3))
    table = pe.relative_attention_bias.weight.detach().cpu()
    bias_ref = orc.t5_bias(table, S, S, bidirectional=False).to(torch.bfloat16).float()
    ref = orc.attn_fwd_bwd(q.permute(0, 2, 1, 3).float(), k.permute(0, 2, 1, 3).float(), v.permute(0, 2, 1, 3).float(),
                           bias_ref, do.permute(0, 2, 1, 3).float(), True, 1.0)
    bk = orc.t5_relative_position_bucket(torch.arange(S)[None, :] - torch.arange(S)[:, None], False)
    dt_ref = torch.zeros(32, H, dtype=torch.float64)
    dt_ref.index_add_(0, bk.reshape(-1), ref[5][0].permute(1, 2, 0).reshape(-1, H))
    mx, rf = orc.error_metrics(dtable, dt_ref)
    assert rf < 1.2e-2, (mx, rf)                                                      # inherits the dBias 16-bit tolerance
