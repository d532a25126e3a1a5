"""Fused multi-tensor AdamWScale step (reference src/utils/adamw_scaled.py).

CPU: the oracle against golden vectors produced by the REFERENCE optimizer itself (oracle/make_golden.py:gen_adamw, three
steps, both of its code paths) -- bit for bit on the per-tensor path, within rounding on the regrouped `_foreach` path --
plus the host logic of the product module (step-size terms with the reference's tensor types, chunk map, descriptor
layout against the C compiler, constructor checks, no CPU fallback).

GPU: the CUDA kernels through the public optimizer against the same golden vectors.  Bars: fp32 parameters 2e-6 relative
to the largest entry of each tensor; bf16 parameters 2 ulp of bf16 (1.6e-2 of the largest entry; the CPU reference does not
fuse multiply-adds and sums the squares in another order, which can flip a 16-bit rounding).  First run on hardware in
round 2 (profiles/r2a_round1_experiments_summary.txt: green; 0.67 of the HBM roof for fp32 parameters).
"""
import ctypes as C
import glob
import os

import numpy as np
import pytest
import torch

import flasht5_b200  # noqa: F401
from flasht5_b200 import _cabi
from flasht5_b200.adamw_scaled import AdamWScale, chunk_map, step_size_terms
from conftest import GOLDEN, ROOT
from oracle import adamw_ref as orc

FILES = sorted(glob.glob(os.path.join(GOLDEN, "adamw_*.npz")))
IDS = [os.path.basename(p)[:-4] for p in FILES]
DT = {"torch.float32": torch.float32, "torch.bfloat16": torch.bfloat16, "torch.float16": torch.float16}


def _t(a, dt=None):
    t = torch.from_numpy(np.asarray(a))
    return t.to(dt) if dt is not None else t


def _case(path):
    z = np.load(path)
    cfg = dict(dt=DT[str(z["dtype"])], kahan=bool(z["kahan"]), foreach=bool(z["foreach"]), wd=float(z["weight_decay"]),
               cb=bool(z["correct_bias"]), lr=float(z["lr"]), b1=float(z["beta1"]), b2=float(z["beta2"]), eps=float(z["eps"]),
               n=int(z["n"]))
    return z, cfg


def test_fixtures_present():
    assert len(FILES) >= 6


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_oracle_matches_reference_golden(path):
    z, c = _case(path)
    for j in range(c["n"]):
        p = _t(z[f"p0_{j}"], c["dt"])
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        comp = torch.zeros_like(p) if (c["kahan"] and c["dt"] != torch.float32) else None
        for step in range(3):
            p, m, v, comp = orc.step_like_reference(p, _t(z[f"g{step}_{j}"], c["dt"]), m, v, comp, step + 1, c["lr"], c["b1"],
                                                    c["b2"], c["eps"], c["wd"], c["cb"])
        for nm, mine in (("p", p), ("m", m), ("v", v)):
            ref = _t(z[f"{nm}{j}"])
            if not c["foreach"]:
                assert torch.equal(mine.float(), ref), (j, nm)
            else:
                tol = 2e-2 if c["dt"] == torch.bfloat16 else 1e-5
                assert (mine.float() - ref).abs().max() <= tol * ref.abs().max() + 1e-12, (j, nm)
        if comp is not None:
            eff_ref = _t(z[f"p{j}"]) + _t(z[f"comp{j}"])
            assert ((p.float() + comp.float()) - eff_ref).abs().max() <= 2e-3 * eff_ref.abs().max()


def test_dtype_faithful_oracle_tracks_the_exact_update():
    g = torch.Generator().manual_seed(0)
    p, gr = torch.randn(500, generator=g), torch.randn(500, generator=g)
    z = torch.zeros(500)
    a = orc.step_like_reference(p, gr, z, z, None, 1, 1e-2, 0.9, 0.999, 1e-6, 0.01)
    b = orc.step_exact(p, gr, z, z, None, 1, 1e-2, 0.9, 0.999, 1e-6, 0.01)
    for x, y in zip(a[:3], b[:3]):
        assert (x.double() - y).abs().max() <= 1e-6 * y.abs().max()
    # Kahan compensation: after many tiny steps the compensated bf16 parameter stays close to the fp64 trajectory,
    # the uncompensated one does not move at all (the update is below half an ulp)
    p16 = torch.ones(64, dtype=torch.bfloat16)
    g16 = torch.full((64,), 1.0, dtype=torch.bfloat16)
    pk, mk, vk, ck = p16.clone(), torch.zeros_like(p16), torch.zeros_like(p16), torch.zeros_like(p16)
    pn, mn, vn = p16.clone(), torch.zeros_like(p16), torch.zeros_like(p16)
    pe, me, ve = p16.double(), torch.zeros(64, dtype=torch.float64), torch.zeros(64, dtype=torch.float64)
    for t in range(1, 41):
        pk, mk, vk, ck = orc.step_like_reference(pk, g16, mk, vk, ck, t, 1e-4, 0.9, 0.999, 1e-6, 0.0)
        pn, mn, vn, _ = orc.step_like_reference(pn, g16, mn, vn, None, t, 1e-4, 0.9, 0.999, 1e-6, 0.0)
        pe, me, ve, _ = orc.step_exact(pe, g16, me, ve, None, t, 1e-4, 0.9, 0.999, 1e-6, 0.0)
    assert torch.equal(pn, p16)
    assert ((pk.double() + ck.double()) - pe).abs().max() < 2e-4 and (pe - 1.0).abs().min() > 3e-3


@pytest.mark.parametrize("cb", [True, False])
def test_step_size_terms_use_the_reference_types(cb):
    for step in (1, 2, 7, 1000):
        base, floor = step_size_terms(step, 2e-2, 0.9, 0.95, cb)
        ref = orc.reference_step_size(step, 2e-2, 0.9, 0.95, cb)
        if cb:
            assert isinstance(ref, torch.Tensor) and ref.dtype == torch.float32
            assert base == float(ref) and floor == float(ref * 1e-3)
        else:
            assert ref == 2e-2 and base == float(np.float32(2e-2)) and floor == 2e-2 * 1e-3


def test_chunk_map_covers_every_element_once():
    numels = (5000, 10, 8192, 1, 4096, 4097)
    first, owner = chunk_map(numels, 4096)
    assert first == [0, 2, 3, 5, 6, 7] and len(owner) == 9
    for i, n in enumerate(numels):
        mine = [c for c, o in enumerate(owner) if o == i]
        assert mine == list(range(first[i], first[i] + (n + 4095) // 4096))       # consecutive, starting at first_chunk
    assert chunk_map((), 4096) == ([], [])


def test_descriptor_layout_matches_c_compiler(tmp_path, lib):
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    fields = [f for f, _ in _cabi.AdamwTensor._fields_]
    body = "".join(f'printf("%zu\\n", offsetof(b200t5_adamw_tensor, {f}));' for f in fields)
    src = tmp_path / "off.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b200t5.h"\nint main(){' + body +
                   'printf("%zu\\n", sizeof(b200t5_adamw_tensor));return 0;}')
    exe = tmp_path / "off"
    subprocess.check_call([cc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out == [getattr(_cabi.AdamwTensor, f).offset for f in fields] + [C.sizeof(_cabi.AdamwTensor)]
    assert lib.b200t5_adamw_chunk_elems() == 4096
    assert lib.b200t5_adamw_workspace_bytes(3, 10) == 256 + 256


def test_entry_point_validates_without_a_gpu(lib):
    buf = (C.c_uint8 * 512)()
    a = (C.addressof(buf) + 255) // 256 * 256
    assert lib.b200t5_adamw_scale_step(None, 0, None, 0, None, 0, 2, 2, 0, 0.9, 0.999, 1e-6, 0, 0, None) == 0     # nothing to do
    assert lib.b200t5_adamw_scale_step(a, 1, a, 1, a, 512, 2, 2, 1, 0.9, 0.999, 1e-6, 0, 0, None) == -1             # Kahan on fp32
    assert "16-bit" in _cabi.last_error()
    assert lib.b200t5_adamw_scale_step(a, 1, a, 1, a, 512, 1, 2, 0, 0.9, 0.999, 1e-6, 0, 0, None) == -2             # bf16 p, fp32 states
    assert lib.b200t5_adamw_scale_step(a, 1, a, 1, a, 512, 2, 2, 0, 1.0, 0.999, 1e-6, 0, 0, None) == -1             # beta out of range
    assert lib.b200t5_adamw_scale_step(a, 1, a, 1, a, 100, 2, 2, 0, 0.9, 0.999, 1e-6, 0, 0, None) == -4             # workspace too small


def test_constructor_mirrors_reference_and_there_is_no_cpu_fallback():
    p = torch.nn.Parameter(torch.zeros(4))
    with pytest.raises(ValueError):
        AdamWScale([p], lr=-1.0)
    with pytest.raises(ValueError):
        AdamWScale([p], betas=(1.0, 0.9))
    with pytest.raises(AssertionError):
        AdamWScale([p], foreach=True, use_state_dtype=torch.bfloat16)
    opt = AdamWScale([p], lr=1e-3, weight_decay=0.1, kahan_sum=True)
    assert opt.defaults == dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.1, foreach=False, kahan_sum=True,
                                correct_bias=True, use_state_dtype=None)
    p.grad = torch.zeros(4)
    with pytest.raises(RuntimeError):
        opt.step()


def _kernel_restated(p, g, m, v, c, step, lr, b1, b2, eps, wd, cb, p_dt, s_dt):
    """csrc/adamw.cu restated with torch fp32 ops (adamw_rms_kernel + adamw_elem, same order, same rounding points;
    multiply-adds left unfused, which moves results by at most an fp32 ulp).  All tensors fp32 holding dtype-representable
    values; returns the same."""
    rnd_p = lambda x: x.to(p_dt).float()   # noqa: E731
    rnd_s = lambda x: x.to(s_dt).float()   # noqa: E731
    f32 = lambda x: float(torch.tensor(x, dtype=torch.float32))   # noqa: E731
    ss_base, ss_floor = step_size_terms(step, lr, b1, b2, cb)
    # pass 1 + 2: sum of squares -> norm -> rms, each a tensor of the parameter dtype
    norm = rnd_p(torch.sqrt((p * p).sum(dtype=torch.float32)))
    rms = rnd_p(norm / f32(p.numel() ** 0.5))
    if float(rms) > f32(1e-3):
        ss = torch.tensor(f32(ss_base)) * rms
        if (not cb) and p_dt != torch.float32:
            ss = rnd_p(ss)
        ss = float(ss)
    else:
        ss = f32(ss_floor)
    value = -ss
    # pass 3
    beta1, beta2, om1, om2, epsf = f32(b1), f32(b2), f32(1.0 - b1), f32(1.0 - b2), f32(eps)
    m = rnd_s(m * beta1)
    m = rnd_s(m + om1 * g)
    v = rnd_s(v * beta2)
    v = rnd_s(v + om2 * (g * g))
    d = rnd_s(torch.sqrt(v))
    d = rnd_s(d + epsf)
    q = value * (m / d)
    if c is not None:
        c = rnd_p(c + q)
        old = p
        p = rnd_p(p + c)
        lost = rnd_p(old - p)
        c = rnd_p(c + lost)
    else:
        p = rnd_p(p + q)
    if wd > 0.0:
        p = rnd_p(p + f32(-lr * wd) * p)
    return p, m, v, c


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_kernel_arithmetic_restated_matches_reference_golden(path):
    """The CUDA kernels' arithmetic, restated line by line on the CPU, against the reference golden vectors at the bars
    of the GPU test below: catches a wrong order of operations or a missing rounding before the kernels ever run."""
    z, c = _case(path)
    tol = 1.6e-2 if c["dt"] == torch.bfloat16 else 2e-6
    for j in range(c["n"]):
        p = _t(z[f"p0_{j}"], c["dt"]).float()
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        comp = torch.zeros_like(p) if (c["kahan"] and c["dt"] != torch.float32) else None
        for step in range(3):
            g = _t(z[f"g{step}_{j}"], c["dt"]).float()
            p, m, v, comp = _kernel_restated(p, g, m, v, comp, step + 1, c["lr"], c["b1"], c["b2"], c["eps"], c["wd"], c["cb"],
                                             c["dt"], c["dt"])
        for nm, mine in (("p", p), ("m", m), ("v", v)):
            ref = _t(z[f"{nm}{j}"])
            err = (mine - ref).abs().max()
            assert err <= tol * ref.abs().max() + 1e-12, (j, nm, float(err))
        if comp is not None:
            eff_ref = _t(z[f"p{j}"]) + _t(z[f"comp{j}"])
            assert ((p + comp) - eff_ref).abs().max() <= 2e-3 * eff_ref.abs().max()


# ------------------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------------------


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_cuda_matches_reference_golden(path):
    z, c = _case(path)
    dev = "cuda:0"
    params = [torch.nn.Parameter(_t(z[f"p0_{j}"], c["dt"]).to(dev)) for j in range(c["n"])]
    opt = AdamWScale(params, lr=c["lr"], betas=(c["b1"], c["b2"]), eps=c["eps"], weight_decay=c["wd"], kahan_sum=c["kahan"],
                     foreach=c["foreach"], correct_bias=c["cb"])
    n0 = _cabi.launch_count()
    for step in range(3):
        for j, prm in enumerate(params):
            prm.grad = _t(z[f"g{step}_{j}"], c["dt"]).to(dev)
        opt.step()
    torch.cuda.synchronize()
    assert _cabi.launch_count() == n0 + 9                                           # three launches per step for ALL tensors
    tol = 1.6e-2 if c["dt"] == torch.bfloat16 else 2e-6
    for j, prm in enumerate(params):
        st = opt.state[prm]
        assert int(st["step"]) == 3
        for nm, mine in (("p", prm.detach()), ("m", st["exp_avg"]), ("v", st["exp_avg_sq"])):
            ref = _t(z[f"{nm}{j}"])
            err = (mine.float().cpu() - ref).abs().max()
            assert err <= tol * ref.abs().max() + 1e-12, (j, nm, float(err))
        if st["kahan_comp"] is not None:
            eff = prm.detach().float().cpu() + st["kahan_comp"].float().cpu()
            eff_ref = _t(z[f"p{j}"]) + _t(z[f"comp{j}"])
            assert (eff - eff_ref).abs().max() <= 2e-3 * eff_ref.abs().max()


@pytest.mark.gpu
def test_cuda_many_tensors_and_state_dict_round_trip():
    """147 tensors of awkward sizes in two parameter groups (decay / no decay), fp32; then a state_dict round trip."""
    dev = "cuda:0"
    g = torch.Generator().manual_seed(1)
    shapes = [(int(torch.randint(1, 9000, (1,), generator=g)),) for _ in range(147)]
    p_cpu = [torch.randn(s, generator=g) for s in shapes]
    g_cpu = [torch.randn(s, generator=g) for s in shapes]
    params = [torch.nn.Parameter(t.to(dev)) for t in p_cpu]
    opt = AdamWScale([{"params": params[:100], "weight_decay": 0.05}, {"params": params[100:], "weight_decay": 0.0}], lr=3e-3)
    for prm, gr in zip(params, g_cpu):
        prm.grad = gr.to(dev)
    opt.step()
    sd = opt.state_dict()
    opt2 = AdamWScale([{"params": params[:100], "weight_decay": 0.05}, {"params": params[100:], "weight_decay": 0.0}], lr=3e-3)
    opt2.load_state_dict(sd)
    opt2.step()
    torch.cuda.synchronize()
    for j, (p0, gr) in enumerate(zip(p_cpu, g_cpu)):
        wd = 0.05 if j < 100 else 0.0
        z = torch.zeros_like(p0)
        p1, m1, v1, _ = orc.step_like_reference(p0, gr, z, z, None, 1, 3e-3, 0.9, 0.999, 1e-6, wd)
        p2, m2, v2, _ = orc.step_like_reference(p1, gr, m1, v1, None, 2, 3e-3, 0.9, 0.999, 1e-6, wd)
        assert (params[j].detach().cpu() - p2).abs().max() <= 2e-6 * p2.abs().max() + 1e-9, j


@pytest.mark.gpu
def test_cuda_mixed_dtype_group_follows_the_group_kahan_flag():
    """Reference :109-113, :127-152: one fp32 parameter resets the group's kahan_sum flag, and then the 16-bit parameters of
    that group -- which own a compensation tensor when they were initialised first -- take the plain update as well
    (ADVICE r1).  The bf16 parameter must end up bit-identical to the same parameter in a group with kahan_sum=False, and
    several steps must not stall on the descriptor upload (pinned staging buffer reused across steps)."""
    dev = "cuda:0"
    g = torch.Generator().manual_seed(7)
    w16 = torch.randn(5000, generator=g).to(torch.bfloat16)
    w32 = torch.randn(3000, generator=g)
    grads16 = [torch.randn(5000, generator=g).to(torch.bfloat16) for _ in range(3)]
    grads32 = [torch.randn(3000, generator=g) for _ in range(3)]

    def run(kahan):
        a, b = torch.nn.Parameter(w16.clone().to(dev)), torch.nn.Parameter(w32.clone().to(dev))
        opt = AdamWScale([a, b], lr=1e-2, kahan_sum=kahan)
        for s in range(3):
            a.grad, b.grad = grads16[s].to(dev), grads32[s].to(dev)
            opt.step()
        torch.cuda.synchronize()
        return a.detach().cpu(), b.detach().cpu(), opt

    a1, b1, opt1 = run(True)
    a0, b0, _ = run(False)
    assert opt1.param_groups[0]["kahan_sum"] is False                       # reset by the fp32 parameter
    assert opt1.state[opt1.param_groups[0]["params"][0]]["kahan_comp"] is not None   # ... after the bf16 one got its tensor
    assert torch.equal(a1, a0) and torch.equal(b1, b0)
