"""CPU check of the FMA-pipe exp2 the developer builds can switch in (csrc/common.cuh: ex2_poly_pair, B200T5_EXP2_POLY):
the device function restated in numpy float32 (same operations in the same order, fused multiply-adds emulated through
float64), with the coefficients parsed from the header so the two cannot drift apart.  Bar: relative error <= 1e-4 over
the whole argument range the kernels produce (t <= 8 forward, t <= ~0 backward), i.e. far inside the 16-bit rounding of P
(bf16 3.9e-3, fp16 4.9e-4)."""
import os
import re

import numpy as np

from conftest import ROOT


def _consts():
    src = open(os.path.join(ROOT, "flasht5_b200", "csrc", "common.cuh")).read()
    get = lambda name: np.float32(float(re.search(r"constexpr float %s = ([-0-9.e]+)f?;" % name, src).group(1)))   # noqa: E731
    return [get("kEx2PolyC%d" % i) for i in range(4)], get("kEx2Magic"), get("kEx2MinT")


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def ex2_poly(t):
    (c0, c1, c2, c3), magic, tmin = _consts()
    t = np.maximum(t.astype(np.float32), tmin)
    tj = (t + magic).astype(np.float32)
    fj = (tj - magic).astype(np.float32)
    f = _fma(fj, np.float32(-1.0) * np.ones_like(t), t)
    p = _fma(np.full_like(t, c3), f, np.full_like(t, c2))
    p = _fma(p, f, np.full_like(t, c1))
    p = _fma(p, f, np.full_like(t, c0))
    bits = p.view(np.uint32) + (tj.view(np.uint32) << np.uint32(23))
    return bits.view(np.float32)


def test_relative_error_over_the_kernel_range():
    rng = np.random.default_rng(0)
    t = np.concatenate([np.linspace(-124.9, 8.0, 400001), rng.uniform(-30, 8, 200000), np.arange(-124, 9) + 0.5,
                        np.arange(-124, 9) - 0.5, np.arange(-124, 9)]).astype(np.float32)
    got = ex2_poly(t).astype(np.float64)
    want = np.exp2(t.astype(np.float64))
    rel = np.abs(got / want - 1.0)
    assert rel.max() < 1e-4, rel.max()
    assert np.all(got > 0) and np.all(np.isfinite(got))


def test_exact_powers_and_monotone_reduction():
    t = np.arange(-100, 9).astype(np.float32)
    got = ex2_poly(t).astype(np.float64)
    assert np.abs(got / np.exp2(t.astype(np.float64)) - 1).max() < 8e-5          # f = 0: only c0's offset remains
    # the reduced argument stays in [-0.5, 0.5] (round-to-nearest magic add), also at the half-way points
    (_, magic, _) = _consts()
    th = (np.arange(-50, 8) + 0.5).astype(np.float32)
    f = th - ((th + magic).astype(np.float32) - magic)
    assert np.all(np.abs(f) <= 0.5)


def test_clamp_below_gives_a_tiny_positive_number_not_garbage():
    t = np.array([-125.0, -126.0, -1000.0, -np.inf], dtype=np.float32)
    got = ex2_poly(t)
    assert np.all(got == got[0]) and 0 < got[0] < 3e-38                           # 2^-125: the caller zeroes empty rows
