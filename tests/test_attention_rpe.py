"""Attention with the T5 relative-position bias computed inside the kernels (second half;
the reference's `fa2_rpe` surface, modeling_flash_t5.py:275-279).

CPU (`-m "not gpu"`): the oracle against golden vectors made from the REFERENCE's dense composition
(oracle/make_golden.py:gen_attn_rpe: RelativePositionalEncoding.compute_bias -> attn_ref -> autograd down to the
embedding table); the host logic (constant ends of the bucket table, band geometry); and a restatement of the
index arithmetic the CUDA kernels use to read the band (forward: one CTA per query block; backward: one CTA per key
block, two 64-column halves, 32-column chunks), checked element by element against the dense bias.

GPU (`-m gpu`): the public op against the same golden vectors and, bit for bit where the arithmetic is identical,
against the dense-bias operator fed with the materialised bias.  Bars as in tests/test_attention_gpu.py
(bf16: O, dV <= 4e-3, dQ, dK <= 1.2e-2 relative Frobenius); the table gradient inherits the dBias bar (1.2e-2).
Both routes of the operator are tested: fused (bias computed in the attention kernels, the default) and composed
(dense producer kernel + dense-bias kernels).
"""
import ctypes as C
import glob
import os

import numpy as np
import pytest
import torch

import flasht5_b200  # noqa: F401
from flasht5_b200 import _cabi
from flasht5_b200 import flash_attention_rpe as rpe
from conftest import GOLDEN, ROOT
from oracle import attn_bias_ref as orc

FILES = sorted(glob.glob(os.path.join(GOLDEN, "rpe_*.npz")))
IDS = [os.path.basename(p)[:-4] for p in FILES]
DEV = "cuda:0"


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _case(path):
    z = np.load(path)
    return z, bool(z["causal"]), float(z["sm_scale"]), int(z["num_buckets"]), int(z["max_distance"])


# ------------------------------------------------------------------------------------------------------------
# CPU: oracle vs reference golden
# ------------------------------------------------------------------------------------------------------------
def test_fixtures_present():
    assert len(FILES) >= 3


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_oracle_matches_reference_golden(path):
    z, causal, scale, nb, maxd = _case(path)
    o, L, dq, dk, dv, dtable = orc.attn_rpe_fwd_bwd(_t(z["q"]), _t(z["k"]), _t(z["v"]), _t(z["table"]), _t(z["do"]), causal,
                                                    scale, num_buckets=nb, max_distance=maxd)
    for name, mine in (("o", o), ("dq", dq), ("dk", dk), ("dv", dv), ("dtable", dtable)):
        mx, rf = orc.error_metrics(mine, _t(z[name]))
        assert rf < 2e-6, (name, mx, rf)


def test_oracle_dtable_is_the_adjoint_of_the_gather():
    """<t5_bias(table), dbias> == <table, t5_dtable(dbias)> for random table / dbias (both bucket kinds)."""
    g = torch.Generator().manual_seed(3)
    for bidir in (True, False):
        table = torch.randn(32, 3, generator=g, dtype=torch.float64)
        dbias = torch.randn(1, 3, 150, 210, generator=g, dtype=torch.float64)
        lhs = (orc.t5_bias(table, 150, 210, bidir) * dbias).sum()
        rhs = (table * orc.t5_dtable(dbias, 150, 210, bidir)).sum()
        assert abs(lhs - rhs) < 1e-9 * (1 + abs(lhs))


# ------------------------------------------------------------------------------------------------------------
# CPU: host logic
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,nb,maxd,bidir", [(1024, 1024, 32, 128, True), (1024, 1024, 32, 128, False),
                                               (70, 330, 16, 64, True), (5, 7, 32, 128, True), (1, 1, 32, 128, True),
                                               (300, 40, 32, 128, False), (4096, 4096, 32, 128, True)])
def test_constant_ends_bracket_every_distinct_bucket(M, N, nb, maxd, bidir):
    lut, zero, lo, hi = rpe.bucket_lut(M, N, nb, maxd, bidir, "cpu")
    assert lut.dtype == torch.int32 and lut.numel() == M + N - 1 and zero == M - 1
    assert lo < hi
    rel = torch.arange(-(M - 1), N)
    if M + N - 1 > 1:
        assert -(M - 1) <= lo and hi <= N - 1
        assert (lut[rel <= lo] == lut[lo + zero]).all()
        assert (lut[rel >= hi] == lut[hi + zero]).all()
        # tight: one step inwards changes the bucket (unless the ends were pulled apart because the whole table is flat)
        if lo + 1 < hi:
            assert lut[lo + 1 + zero] != lut[lo + zero] or lut[hi - 1 + zero] != lut[hi + zero]
    assert rpe.band_len(lo, hi) == hi - lo + 511


def test_t5_tables_have_a_short_band_at_any_sequence_length():
    for S in (512, 1024, 4096, 16384):
        for bidir in (True, False):
            _, _, lo, hi = rpe.bucket_lut(S, S, 32, 128, bidir, "cpu")
            assert hi - lo <= 256 and rpe.band_len(lo, hi) <= 767 <= rpe.MAX_BAND_LEN


def test_band_len_agrees_with_the_library(lib):
    for lo, hi in ((-91, 91), (-113, 0), (-4, 6), (0, 1)):
        assert lib.b200t5_rpe_band_len(lo, hi) == rpe.band_len(lo, hi)
    assert lib.b200t5_rpe_band_len(5, 5) == 0


def _band_reference(table, lut, zero, lo, hi):
    """What b200t5_rpe_band writes (t5_bias.cu:rpe_band_kernel), restated: (H, band_len) over rel = lo-255 .. hi+255."""
    band_lo = lo - rpe.BAND_PAD
    idx = (torch.arange(rpe.band_len(lo, hi)) + band_lo + zero).clamp(0, lut.numel() - 1)
    return table[lut[idx].long()].t().contiguous(), band_lo


def _fwd_kernel_bias(band_row, band_lo, lo, hi, M, N):
    """Index arithmetic of attn_fwd.cu (bias mode 3): CTA = 128-row query block, tiles of 128 keys, thread = row."""
    Mp, Np = -(-M // 128) * 128, -(-N // 128) * 128
    out = torch.empty(Mp, Np, dtype=band_row.dtype)
    c = torch.arange(128)
    for row0 in range(0, Mp, 128):
        for col0 in range(0, Np, 128):
            rel_min, rel_max = col0 - row0 - 127, col0 - row0 + 127
            if rel_max <= lo or rel_min >= hi:
                out[row0:row0 + 128, col0:col0 + 128] = band_row[(lo if rel_max <= lo else hi) - band_lo]
            else:
                for r in range(128):
                    base = col0 - (row0 + r) - band_lo
                    assert 0 <= base and base + 127 < band_row.numel(), "general tile left the band"
                    out[row0 + r, col0:col0 + 128] = band_row[base + c]
    return out[:M, :N]


def _bwd_kernel_bias(band_row, band_lo, lo, hi, M, N):
    """Index arithmetic of attn_bwd_v3.cu / attn_bwd.cu (bias mode 3): CTA = 128-key block, loop over query blocks,
    warpgroup wg owns 64 columns, chunks of 32 (v2) -- the 8-column groups of attn_bwd.cu index the same addresses."""
    Mp, Np = -(-M // 128) * 128, -(-N // 128) * 128
    out = torch.empty(Mp, Np, dtype=band_row.dtype)
    e = torch.arange(32)
    for col0 in range(0, Np, 128):
        for mrow0 in range(0, Mp, 128):
            rel_min, rel_max = col0 - mrow0 - 127, col0 - mrow0 + 127
            const = rel_max <= lo or rel_min >= hi
            cval = band_row[(lo if rel_max <= lo else hi) - band_lo] if const else None
            for wg in range(2):
                for ch in range(2):
                    cs = col0 + wg * 64 + ch * 32
                    if const:
                        out[mrow0:mrow0 + 128, cs:cs + 32] = cval
                    else:
                        for r in range(128):
                            base = col0 + wg * 64 + ch * 32 - (mrow0 + r) - band_lo
                            assert 0 <= base and base + 31 < band_row.numel()
                            out[mrow0 + r, cs:cs + 32] = band_row[base + e]
    return out[:M, :N]


@pytest.mark.parametrize("M,N,nb,maxd,bidir", [(300, 300, 32, 128, True), (200, 200, 32, 128, False),
                                               (70, 330, 16, 64, True), (330, 70, 32, 128, True), (129, 640, 32, 128, False),
                                               (384, 384, 32, 16, True), (5, 7, 32, 128, True)])
def test_kernel_band_indexing_reproduces_the_dense_bias(M, N, nb, maxd, bidir):
    g = torch.Generator().manual_seed(M * 7 + N)
    table = torch.randn(nb, 2, generator=g)
    lut, zero, lo, hi = rpe.bucket_lut(M, N, nb, maxd, bidir, "cpu")
    band, band_lo = _band_reference(table, lut, zero, lo, hi)
    dense = orc.t5_bias(table, M, N, bidir, nb, maxd)[0]                       # (H, M, N)
    for h in range(2):
        assert torch.equal(_fwd_kernel_bias(band[h], band_lo, lo, hi, M, N), dense[h])
        assert torch.equal(_bwd_kernel_bias(band[h], band_lo, lo, hi, M, N), dense[h])


def _band_reducer(ws, lut, zero, lo, hi, M, N, nb, causal):
    """Index arithmetic of t5_bias.cu:rpe_dtable_band_kernel restated: ws (G, H, M, N) dS surface -> dtable (nb, H).  Thread
    t of a CTA walks the wrapped diagonal w = t % 128 over its half of the rows; a wrapped diagonal holds the true diagonals
    d = w (c = r + w < 128) and d = w - 128."""
    G, H = ws.shape[:2]
    dtable = torch.zeros(nb, H, dtype=torch.float64)
    wsum = ws.sum(0).numpy()                                   # the CTAs of the G slices add into the same table
    lut_l = lut.tolist()
    for row0 in range(0, M, 128):
        for col0 in range(0, N, 128):
            rel_min, rel_max = col0 - row0 - 127, col0 - row0 + 127
            if rel_max <= lo or rel_min >= hi:
                continue
            if causal and col0 > row0 + 127 + (N - M):
                continue
            for h in range(H):
                sdiag = [0.0] * 256
                tile = wsum[h]
                for t in range(256):
                    w, half = t & 127, t >> 7
                    for rr in range(64):
                        r = half * 64 + rr
                        c = r + w
                        wrapped = c >= 128
                        c &= 127
                        m, n = row0 + r, col0 + c
                        if m < M and n < N:
                            sdiag[w if wrapped else w + 128] += float(tile[m, n])
                for t in range(256):
                    idx = min(max(col0 - row0 + (t - 128) + zero, 0), len(lut_l) - 1)
                    dtable[lut_l[idx], h] += sdiag[t]
    return dtable


@pytest.mark.parametrize("M,N,bidir,causal,maxd", [(300, 300, True, False, 16), (130, 400, True, False, 16),
                                                    (390, 390, False, True, 32)])
def test_band_reducer_indexing_matches_the_scatter(M, N, bidir, causal, maxd):
    """The developer path that reduces the table gradient straight from the non-constant tiles of the dS surface: its
    diagonal walk, restated, must give the oracle's scatter-add over exactly those tiles."""
    g = torch.Generator().manual_seed(M + N)
    lut, zero, lo, hi = rpe.bucket_lut(M, N, 32, maxd, bidir, "cpu")        # small max_distance: constant tiles exist at these sizes
    ws = torch.randn(2, 2, M, N, generator=g, dtype=torch.float64)
    want_in = torch.zeros(1, 2, M, N, dtype=torch.float64)
    for row0 in range(0, M, 128):
        for col0 in range(0, N, 128):
            rel_min, rel_max = col0 - row0 - 127, col0 - row0 + 127
            const = rel_max <= lo or rel_min >= hi
            masked = causal and col0 > row0 + 127 + (N - M)
            if not const and not masked:
                want_in[0, :, row0:row0 + 128, col0:col0 + 128] = ws[:, :, row0:row0 + 128, col0:col0 + 128].sum(0)
    assert (want_in == 0).any()                                           # some tiles are constant (or masked) at these sizes
    want = orc.t5_dtable(want_in, M, N, bidir, 32, maxd)
    got = _band_reducer(ws, lut, zero, lo, hi, M, N, 32, causal)
    assert torch.allclose(got, want, atol=1e-9)


def test_rpe_struct_layout_matches_c_compiler(tmp_path):
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    fields = [f for f, _ in _cabi.RpeParams._fields_]
    body = "".join(f'printf("%zu\\n", offsetof(b200t5_rpe_params, {f}));' for f in fields)
    src = tmp_path / "off.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b200t5.h"\nint main(){' + body +
                   'printf("%zu\\n", sizeof(b200t5_rpe_params));return 0;}')
    exe = tmp_path / "off"
    subprocess.check_call([cc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out == [getattr(_cabi.RpeParams, f).offset for f in fields] + [C.sizeof(_cabi.RpeParams)]


def test_public_surface_and_no_cpu_fallback():
    from flasht5_b200 import flash_attention_v2_rpe, FlashAttentionRPE   # noqa: F401
    q = torch.zeros(1, 2, 16, 64, dtype=torch.bfloat16)
    w = torch.zeros(2, 32)
    with pytest.raises(RuntimeError):
        flash_attention_v2_rpe(q, q, q, w, 128)
    with pytest.raises(ValueError):
        flash_attention_v2_rpe(q, q, q, torch.zeros(3, 32), 128)


def test_fake_impls_shape_contract():
    """Meta-device contract of the three new ops (what torch.compile traces)."""
    B, H, M, N, D = 2, 3, 40, 56, 64
    q = torch.empty(B, H, M, D, dtype=torch.bfloat16, device="meta")
    k = torch.empty(B, H, N, D, dtype=torch.bfloat16, device="meta")
    table = torch.empty(32, H, device="meta")
    lut = torch.empty(M + N - 1, dtype=torch.int32, device="meta")
    band = torch.ops.b200t5.rpe_band(table, lut, M - 1, -20, 30, torch.bfloat16)
    assert band.shape == (H, 30 + 20 + 511) and band.dtype == torch.float32
    o, L = torch.ops.b200t5.attn_rpe_fwd(q, k, k, band, -20, 30, False, 1.0)
    assert o.shape == q.shape and o.dtype == q.dtype and L.shape == (B, H, M) and L.dtype == torch.float32
    dq, dk, dv, dt = torch.ops.b200t5.attn_rpe_bwd(o, o, q, k, k, band, lut, M - 1, -20, 30, 32, L, False, 1.0)
    assert dq.shape == q.shape and dk.shape == k.shape and dv.shape == k.shape
    assert dt.shape == (32, H) and dt.dtype == torch.float32


def test_entry_points_validate_without_a_gpu(lib):
    """Argument errors are reported before any CUDA call."""
    p = _cabi.AttnParams()
    r = _cabi.RpeParams()
    assert lib.b200t5_attn_rpe_fwd(C.byref(p), None) < 0
    p.B = p.H = 1
    p.M = p.N = 128
    p.D = 64
    p.dtype = _cabi.BF16
    buf = (C.c_uint8 * 64)()
    addr = (C.addressof(buf) + 15) // 16 * 16
    for f in ("q", "k", "v", "o"):
        setattr(p, f, addr)
        setattr(p, f + "_strides", _cabi.I64x4(8192, 8192, 64, 1))
    p.lse = addr
    r.const_lo, r.const_hi = 10, 10
    assert lib.b200t5_attn_rpe_fwd(C.byref(p), C.byref(r)) == -1 and "const_lo" in _cabi.last_error()
    r.const_lo, r.const_hi = -20000, 20000
    assert lib.b200t5_attn_rpe_fwd(C.byref(p), C.byref(r)) == -2                  # band too long: unsupported
    r.const_lo, r.const_hi = -91, 91
    assert lib.b200t5_attn_rpe_fwd(C.byref(p), C.byref(r)) == -1 and "band" in _cabi.last_error()
    p.bias = addr
    p.bias_B = p.bias_H = 1
    r.band = addr
    assert lib.b200t5_attn_rpe_fwd(C.byref(p), C.byref(r)) == -1 and "bias == NULL" in _cabi.last_error()


# ------------------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------------------
TOL = {"o": 4e-3, "dv": 4e-3, "dq": 1.2e-2, "dk": 1.2e-2, "dtable": 1.2e-2}
FUSED_MODES = [False, True]


def _run_rpe(z, causal, scale, maxd, fused, dtype=torch.bfloat16):
    from flasht5_b200 import flash_attention_v2_rpe
    q, k, v, do = (_t(z[n]).to(dtype).to(DEV) for n in ("q", "k", "v", "do"))
    w = _t(z["table"]).t().contiguous().to(DEV).requires_grad_(True)          # (H, num_buckets), as the reference passes it
    q.requires_grad_(True), k.requires_grad_(True), v.requires_grad_(True)
    o = flash_attention_v2_rpe(q, k, v, w, maxd, causal=causal, sm_scale=scale, fused=fused)
    dq, dk, dv, dw = torch.autograd.grad(o, (q, k, v, w), do)
    torch.cuda.synchronize()
    return {"o": o, "dq": dq, "dk": dk, "dv": dv, "dtable": dw.t()}


@pytest.mark.gpu
@pytest.mark.parametrize("fused", FUSED_MODES, ids=lambda f: "fused" if f else "composed")
@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_cuda_matches_reference_golden(path, fused):
    z, causal, scale, nb, maxd = _case(path)
    got = _run_rpe(z, causal, scale, maxd, fused)
    for name, tol in TOL.items():
        mx, rf = orc.error_metrics(got[name], _t(z[name]))
        assert rf < tol, (name, mx, rf)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 4, 512, 512, 64, False), (2, 4, 512, 512, 64, True), (1, 2, 300, 700, 128, False),
                                   (2, 2, 640, 384, 32, False), (2, 2, 384, 640, 32, True), (1, 3, 130, 130, 16, False), (3, 8, 1024, 1024, 64, False)],
                         ids=lambda s: "x".join(map(str, s)))
def test_cuda_fused_equals_dense_operator(shape):
    """Same 16-bit bias values, same arithmetic: O and LSE-dependent outputs are bit-identical to the dense-bias
    operator; dQ (and dTable) only differ by the order of 16-bit partial sums.
    (No causal case with M > N: there every visible position lies beyond max_distance, i.e. in ONE bucket, and the rows of dS
    sum to zero -- the table gradient is rounding noise around 0 on both routes and a relative comparison says nothing.)"""
    from flasht5_b200 import flash_attention_v2_rpe
    B, H, M, N, D, causal = shape
    g = torch.Generator().manual_seed(11)
    mk = lambda s: torch.randn(B, s, H, D, generator=g).to(torch.bfloat16).to(DEV).permute(0, 2, 1, 3)   # noqa: E731
    q, k, v, do = mk(M), mk(N), mk(N), mk(M)
    w = (0.5 * torch.randn(H, 32, generator=g)).to(DEV)
    outs = {}
    for fused in (False, True):
        qq, kk, vv, ww = (t.detach().clone().requires_grad_(True) for t in (q, k, v, w))
        o = flash_attention_v2_rpe(qq, kk, vv, ww, 128, causal=causal, sm_scale=1.0, fused=fused)
        outs[fused] = (o,) + torch.autograd.grad(o, (qq, kk, vv, ww), do)
    torch.cuda.synchronize()
    for i, name in enumerate(("o", "dq", "dk", "dv", "dw")):
        a, b = outs[False][i], outs[True][i]
        if name in ("o", "dk", "dv"):
            assert torch.equal(a, b), name
        else:
            mx, rf = orc.error_metrics(b, a.double())
            assert rf < 4e-3, (name, mx, rf)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 4, 512, 512, 64, False), (2, 2, 640, 384, 16, True)], ids=lambda s: "x".join(map(str, s)))
def test_cuda_constant_tile_skip(shape):
    """Backward of the fused operator (compile-time level 2, csrc/api.cu): sub-tiles entirely beyond a constant end of the
    bucket table keep their dS in the kernel (two fp32 sums per CTA), the other tiles reduce into the transposed surface and
    the table gradient is folded straight from it.  dQ/dK/dV must equal the composed dense route (to the order of the 16-bit
    partial sums for dQ); the table gradient is compared with the fp64 oracle on the scale of the |dS| mass that flows into
    each bucket (the second shape is the degenerate one: every visible position falls into ONE bucket, whose exact gradient
    is 0 because rows of dS sum to zero)."""
    from flasht5_b200 import flash_attention_v2_rpe
    B, H, M, N, D, causal = shape
    g = torch.Generator().manual_seed(11)
    mk = lambda s: torch.randn(B, s, H, D, generator=g).to(torch.bfloat16).permute(0, 2, 1, 3)   # noqa: E731
    q, k, v, do = mk(M), mk(N), mk(N), mk(M)
    w = 0.5 * torch.randn(H, 32, generator=g)
    outs = {}
    levels = ("dense", "fused")
    for route in levels:
        qq, kk, vv, ww = (t.to(DEV).requires_grad_(True) for t in (q, k, v, w))
        o = flash_attention_v2_rpe(qq, kk, vv, ww, 128, causal=causal, sm_scale=1.0, fused=(route == "fused"))
        outs[route] = (o,) + torch.autograd.grad(o, (qq, kk, vv, ww), do.to(DEV))
        torch.cuda.synchronize()
    for i, name in enumerate(("o", "dq", "dk", "dv")):
        a, b = outs["dense"][i], outs["fused"][i]
        if name == "dq":                                  # order of 16-bit partial sums at L2 differs run to run (~2.5e-3)
            assert orc.error_metrics(b, a.double())[1] < 6e-3
        else:
            assert torch.equal(a, b), name
    table = w.t().contiguous()
    bias = orc.t5_bias(table, M, N, bidirectional=not causal).to(torch.bfloat16).float()
    ref = orc.attn_fwd_bwd(q.float(), k.float(), v.float(), bias, do.float(), causal, 1.0)
    dt_ref = orc.t5_dtable(ref[5], M, N, not causal)
    mass = orc.t5_dtable(ref[5].abs(), M, N, not causal)                       # |dS| flowing into each bucket
    for route in levels:
        err = (outs[route][4].t().double().cpu() - dt_ref).norm() / mass.norm()
        assert err < 4e-3, (route, float(err))


@pytest.mark.gpu
def test_cuda_library_launches_for_rpe():
    from flasht5_b200 import flash_attention_v2_rpe
    n0 = _cabi.launch_count()
    q = torch.randn(1, 2, 128, 64, device=DEV, dtype=torch.bfloat16)
    w = torch.randn(2, 32, device=DEV)
    flash_attention_v2_rpe(q, q, q, w, 128)
    torch.cuda.synchronize()
    assert _cabi.launch_count() == n0 + 2                                     # band + attention forward
