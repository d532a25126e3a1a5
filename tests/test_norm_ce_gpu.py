"""GPU parity: RMSNorm and cross-entropy(+z-loss) CUDA kernels against the reference golden
vectors and the fp64 oracle on seeded inputs.  Tolerances follow the reference's own tests
(tests/layer_norm_triton/test_layer_norm.py:22-43 and tests/cross_entropy_triton/test_cross_entropy.py:27-49:
allclose(atol=1e-2, rtol=0)); fp32 I/O is held to 1e-5 relative."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import attn_bias_ref as orc
from oracle import ce_ref, rmsnorm_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    return torch.from_numpy(np.asarray(a))


RMS_FILES = sorted(glob.glob(os.path.join(GOLDEN, "rmsnorm_*.npz")))
CE_FILES = sorted(glob.glob(os.path.join(GOLDEN, "ce_*.npz")))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("path", RMS_FILES, ids=[os.path.basename(p)[:-4] for p in RMS_FILES])
def test_rmsnorm_golden(path, dtype):
    from flasht5_b200 import fast_rms_layernorm
    z = np.load(path)
    x = _t(z["x"]).to(DEV, dtype).requires_grad_(True)
    w = _t(z["w"]).to(DEV, dtype).requires_grad_(True)
    dy = _t(z["dy"]).to(DEV, dtype)
    y = fast_rms_layernorm(x, w, float(z["eps"]))
    dx, dw = torch.autograd.grad(y, (x, w), dy)
    assert y.dtype == dtype and dx.dtype == dtype and dw.dtype == dtype
    if dtype == torch.float32:
        for name, mine in (("y", y), ("dx", dx), ("dw", dw)):
            mx, rf = orc.error_metrics(mine, _t(z[name]))
            assert rf < 1e-5, (name, mx, rf)
    else:
        atol = 1e-2 if dtype == torch.float16 else 6e-2      # bf16: 8 mantissa bits on |y| up to ~8
        for name, mine in (("y", y), ("dx", dx)):
            mx, rf = orc.error_metrics(mine, _t(z[name]))
            assert mx <= atol and rf < 5e-3, (name, mx, rf)
        mx, rf = orc.error_metrics(dw, _t(z["dw"]))
        assert rf < 8e-3, ("dw", mx, rf)


@pytest.mark.parametrize("rows,n,dtype,wdtype", [
    (4096, 512, torch.bfloat16, torch.bfloat16), (1000, 768, torch.bfloat16, torch.float32),
    (333, 1024, torch.float16, torch.float16), (64, 2048, torch.bfloat16, torch.bfloat16),
    (17, 4096, torch.float32, torch.float32), (9, 8192, torch.bfloat16, torch.bfloat16),
    (5, 1001, torch.bfloat16, torch.bfloat16), (3, 10000, torch.float32, torch.float32), (1, 64, torch.float16, torch.float32),
])
def test_rmsnorm_seeded_vs_oracle(rows, n, dtype, wdtype):
    from flasht5_b200 import fast_rms_layernorm
    g = torch.Generator().manual_seed(rows * 131 + n)
    x = (torch.randn(rows, n, generator=g) * 1.5).to(dtype)
    w = (1 + 0.1 * torch.randn(n, generator=g)).to(wdtype)
    dy = torch.randn(rows, n, generator=g).to(dtype)
    y_ref, rstd = rmsnorm_ref.rmsnorm_fwd(x, w, 1e-6)
    dx_ref, dw_ref = rmsnorm_ref.rmsnorm_bwd(dy, x, w, rstd)
    xd, wd = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True)
    y = fast_rms_layernorm(xd, wd, 1e-6)
    dx, dw = torch.autograd.grad(y, (xd, wd), dy.to(DEV))
    tol = 2e-6 if dtype == torch.float32 else (6e-4 if dtype == torch.float16 else 5e-3)
    for name, mine, ref in (("y", y, y_ref), ("dx", dx, dx_ref)):
        mx, rf = orc.error_metrics(mine, ref)
        assert rf < tol, (name, mx, rf)
    tolw = 2e-5 if wdtype == torch.float32 else (1e-3 if wdtype == torch.float16 else 6e-3)
    mx, rf = orc.error_metrics(dw, dw_ref)
    assert rf < tolw, ("dw", mx, rf)


def test_rmsnorm_3d_input_and_row_stride():
    from flasht5_b200 import fast_rms_layernorm
    g = torch.Generator().manual_seed(5)
    big = torch.randn(6, 50, 1024, generator=g).to(torch.bfloat16).to(DEV)
    x = big[:, :, :512]                                    # row stride 1024, unit last stride
    w = torch.ones(512, dtype=torch.bfloat16, device=DEV)
    y = fast_rms_layernorm(x, w, 1e-6)
    assert y.shape == x.shape
    y_ref, _ = rmsnorm_ref.rmsnorm_fwd(x.float().cpu().reshape(-1, 512), w.float().cpu(), 1e-6)
    mx, rf = orc.error_metrics(y.reshape(-1, 512), y_ref)
    assert rf < 5e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("path", CE_FILES, ids=[os.path.basename(p)[:-4] for p in CE_FILES])
def test_ce_golden(path, dtype):
    from flasht5_b200 import cross_entropy_loss
    z = np.load(path)
    logits = _t(z["logits"]).to(DEV, dtype).requires_grad_(True)
    labels = _t(z["labels"]).to(DEV)
    zl, sm = float(z["z"]), float(z["smoothing"])
    losses, z_losses = cross_entropy_loss(logits, labels, label_smoothing=sm, lse_square_scale=zl)
    n_valid = int((labels != -100).sum())
    loss = losses.sum() / n_valid                          # the golden is the mean over non-ignored rows
    (dlogits,) = torch.autograd.grad(loss, logits)
    assert losses.dtype == torch.float32 and dlogits.dtype == dtype
    assert abs(loss.item() - float(z["loss_mean_valid"])) < (1e-4 if dtype == torch.float32 else 1e-2)
    assert torch.all(losses[labels == -100] == 0) and torch.all(dlogits[labels == -100] == 0)
    assert torch.allclose(dlogits.float().cpu(), _t(z["dlogits"]), atol=1e-2 if dtype != torch.float32 else 1e-6, rtol=0)
    mx, rf = orc.error_metrics(dlogits, _t(z["dlogits"]))
    assert rf < (1e-5 if dtype == torch.float32 else 6e-3), (mx, rf)


@pytest.mark.parametrize("rows,V,dtype,z,sm,scale,inplace", [
    (512, 32768, torch.bfloat16, 1e-4, 0.0, 1.0, False), (64, 32128, torch.float16, 0.0, 0.1, 1.0, False),
    (33, 32102, torch.float32, 2.0, 0.1, 1.0, True), (7, 50000, torch.bfloat16, 1.0, 0.0, 0.5, True),
    (5, 999, torch.float32, 1e-4, 0.05, 1.3, False), (3, 8, torch.bfloat16, 0.0, 0.0, 1.0, False),
])
def test_ce_seeded_vs_oracle(rows, V, dtype, z, sm, scale, inplace):
    from flasht5_b200 import cross_entropy_loss
    g = torch.Generator().manual_seed(rows + V)
    logits = (torch.randn(rows, V, generator=g) * 2.5).to(dtype)
    labels = torch.randint(0, V, (rows,), generator=g)
    labels[rows // 2] = -100
    dl = torch.randn(rows, generator=g)
    l_ref, z_ref, lse_ref = ce_ref.ce_fwd(logits, labels, sm, scale, z)
    d_ref = ce_ref.ce_bwd(dl, logits, lse_ref, labels, sm, scale, z)
    ld = logits.to(DEV).requires_grad_(True)
    src = ld.clone() if inplace else ld                     # in-place backward may not clobber a leaf
    losses, z_losses = cross_entropy_loss(src, labels.to(DEV), None, sm, scale, z, -100, inplace)
    (dlogits,) = torch.autograd.grad(losses, ld, dl.to(DEV))
    tol = 1e-5 if dtype == torch.float32 else 1e-3
    assert torch.allclose(losses.double().cpu(), l_ref, atol=tol * max(1.0, l_ref.abs().max().item()), rtol=tol)
    assert torch.allclose(z_losses.double().cpu(), z_ref, atol=tol * max(1.0, z_ref.abs().max().item()), rtol=tol)
    mx, rf = orc.error_metrics(dlogits, d_ref)
    assert rf < (1e-5 if dtype == torch.float32 else (8e-4 if dtype == torch.float16 else 6e-3)), (mx, rf)


def test_ce_precomputed_lse_and_all_rows_mean():
    from flasht5_b200 import cross_entropy_loss
    g = torch.Generator().manual_seed(9)
    logits = torch.randn(16, 4096, generator=g).to(DEV)
    labels = torch.randint(0, 4096, (16,), generator=g).to(DEV)
    labels[3] = -100
    l0, z0 = cross_entropy_loss(logits, labels, lse_square_scale=1e-4)
    lse = torch.logsumexp(logits, -1)
    l1, z1 = cross_entropy_loss(logits, labels, precomputed_lse=lse, lse_square_scale=1e-4)
    assert torch.allclose(l0, l1, atol=1e-5) and torch.allclose(z0, z1, atol=1e-7)
    # the model takes .mean() over ALL rows including ignored ones (modeling_flash_t5.py:64-68)
    assert l0[3] == 0 and abs(l0.mean().item() - l0.sum().item() / 16) < 1e-6
