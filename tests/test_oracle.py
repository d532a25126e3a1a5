"""CPU: the oracle restatements against the golden vectors generated from the REFERENCE itself
(oracle/make_golden.py imported /root/reference in the build container) and against the
known-answer bucket vector recorded in SURVEY.md section 8c."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import attn_bias_ref as orc
from oracle import ce_ref, rmsnorm_ref
from conftest import GOLDEN


def _t(a):
    return torch.from_numpy(np.asarray(a))


ATTN_FILES = sorted(glob.glob(os.path.join(GOLDEN, "attn_*.npz")))


def test_golden_files_present():
    assert len(ATTN_FILES) >= 6
    assert len(glob.glob(os.path.join(GOLDEN, "rmsnorm_*.npz"))) >= 3
    assert len(glob.glob(os.path.join(GOLDEN, "ce_*.npz"))) >= 3


@pytest.mark.parametrize("path", ATTN_FILES, ids=[os.path.basename(p)[:-4] for p in ATTN_FILES])
def test_attention_oracle_matches_reference_golden(path):
    z = np.load(path)
    q, k, v, do = (_t(z[n]) for n in ("q", "k", "v", "do"))
    bias = _t(z["bias"]) if "bias" in z.files else None
    causal = bool(z["causal"])
    scale = None if np.isnan(z["sm_scale"]) else float(z["sm_scale"])
    valid = _t(z["valid_rows"])
    do_m = torch.where(valid.view(1, 1, -1, 1), do, torch.zeros_like(do))
    o, L, dq, dk, dv, db = orc.attn_fwd_bwd(q, k, v, bias, do_m, causal, scale)
    for name, mine in (("o", o), ("dq", dq), ("dk", dk), ("dv", dv)):
        mx, rf = orc.error_metrics(mine, _t(z[name]))
        assert rf < 2e-6, (name, mx, rf)          # reference ran in fp32, oracle in fp64
    if bias is not None:
        mx, rf = orc.error_metrics(db, _t(z["dbias"]))
        assert rf < 2e-6, ("dbias", mx, rf)
        assert db.shape == bias.shape
    # LSE: -inf exactly on rows with no visible key, finite elsewhere
    assert torch.equal(torch.isinf(L).any(dim=(0, 1)), ~valid)
    # empty rows produce O = 0 (reference kernel :470-473)
    assert torch.all(o[:, :, ~valid] == 0)


def test_t5_bucket_known_answers():
    z = np.load(os.path.join(GOLDEN, "t5_buckets.npz"))
    rel = _t(z["rel"])
    assert orc.t5_relative_position_bucket(rel.clone(), True).tolist() == z["bidirectional"].tolist()
    assert orc.t5_relative_position_bucket(rel.clone(), False).tolist() == z["unidirectional"].tolist()
    # SURVEY.md section 8c, recorded from the reference in the survey container
    assert z["bidirectional"].tolist() == [15, 15, 15, 14, 12, 10, 8, 8, 7, 1, 0, 17, 23, 24, 24, 26, 28, 30, 31, 31, 31]
    assert z["unidirectional"].tolist() == [31, 31, 31, 26, 21, 16, 9, 8, 7, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]


def test_t5_bias_is_toeplitz():
    table = torch.randn(32, 4, generator=torch.Generator().manual_seed(0))
    b = orc.t5_bias(table, 40, 56, bidirectional=True)
    assert b.shape == (1, 4, 40, 56)
    assert torch.equal(b[0, :, 1:, 1:], b[0, :, :-1, :-1])


def test_attention_oracle_lse_and_softmax_identities():
    g = torch.Generator().manual_seed(3)
    q, k, v = (torch.randn(2, 2, 24, 16, generator=g) for _ in range(3))
    bias = torch.randn(1, 2, 24, 24, generator=g)
    o, L = orc.attn_fwd(q, k, v, bias, causal=True, sm_scale=0.3)
    s = torch.einsum("bhmd,bhnd->bhmn", q.double(), k.double()) * 0.3 + bias.double()
    s = s.masked_fill(torch.triu(torch.ones(24, 24, dtype=torch.bool), 1), float("-inf"))
    assert torch.allclose(L, torch.logsumexp(s, -1), atol=1e-12)
    assert torch.allclose(o, torch.softmax(s, -1) @ v.double(), atol=1e-12)
    # V = ones -> O = ones on every row with a visible key
    o1, _ = orc.attn_fwd(q, k, torch.ones_like(v), bias, causal=True, sm_scale=0.3)
    assert torch.allclose(o1, torch.ones_like(o1), atol=1e-12)


def test_attention_oracle_backward_matches_autograd():
    g = torch.Generator().manual_seed(4)
    B, H, M, N, D = 2, 3, 20, 28, 16
    q = torch.randn(B, H, M, D, generator=g, dtype=torch.float64, requires_grad=True)
    k = torch.randn(B, H, N, D, generator=g, dtype=torch.float64, requires_grad=True)
    v = torch.randn(B, H, N, D, generator=g, dtype=torch.float64, requires_grad=True)
    for shape in ((1, H, M, N), (B, 1, M, N), (1, 1, M, N), (B, H, M, N)):
        bias = torch.randn(*shape, generator=g, dtype=torch.float64, requires_grad=True)
        do = torch.randn(B, H, M, D, generator=g, dtype=torch.float64)
        s = q @ k.transpose(2, 3) * 0.7 + bias
        mask = torch.arange(M).unsqueeze(-1) + (N - M) >= torch.arange(N)
        s = s.masked_fill(~mask, float("-inf"))
        o_ag = torch.softmax(s, -1) @ v
        gq, gk, gv, gb = torch.autograd.grad(o_ag, (q, k, v, bias), do)
        o, L, dq, dk, dv, db = orc.attn_fwd_bwd(q.detach(), k.detach(), v.detach(), bias.detach(), do, True, 0.7)
        for a, b_ in ((o, o_ag), (dq, gq), (dk, gk), (dv, gv), (db, gb)):
            assert torch.allclose(a, b_.detach(), atol=1e-10), shape


RMS_FILES = sorted(glob.glob(os.path.join(GOLDEN, "rmsnorm_*.npz")))
CE_FILES = sorted(glob.glob(os.path.join(GOLDEN, "ce_*.npz")))


@pytest.mark.parametrize("path", RMS_FILES, ids=[os.path.basename(p)[:-4] for p in RMS_FILES])
def test_rmsnorm_oracle_matches_reference_golden(path):
    z = np.load(path)
    x, w, dy = _t(z["x"]), _t(z["w"]), _t(z["dy"])
    y, rstd = rmsnorm_ref.rmsnorm_fwd(x, w, float(z["eps"]))
    dx, dw = rmsnorm_ref.rmsnorm_bwd(dy, x, w, rstd)
    for name, mine in (("y", y), ("dx", dx), ("dw", dw)):
        mx, rf = orc.error_metrics(mine, _t(z[name]))
        assert rf < 2e-6, (name, mx, rf)


@pytest.mark.parametrize("path", CE_FILES, ids=[os.path.basename(p)[:-4] for p in CE_FILES])
def test_ce_oracle_matches_reference_golden(path):
    z = np.load(path)
    logits, labels = _t(z["logits"]), _t(z["labels"])
    zl, sm = float(z["z"]), float(z["smoothing"])
    losses, z_losses, lse = ce_ref.ce_fwd(logits, labels, sm, 1.0, zl)
    n_valid = int((labels != -100).sum())
    assert abs(losses.sum().item() / n_valid - float(z["loss_mean_valid"])) < 2e-5 * max(1.0, abs(float(z["loss_mean_valid"])))
    assert torch.all(losses[labels == -100] == 0) and torch.all(z_losses[labels == -100] == 0)
    dl = torch.full((logits.shape[0],), 1.0 / n_valid, dtype=torch.float64)
    dlogits = ce_ref.ce_bwd(dl, logits, lse, labels, sm, 1.0, zl)
    mx, rf = orc.error_metrics(dlogits, _t(z["dlogits"]))
    assert rf < 5e-6, (mx, rf)
