"""Generate tests/golden/*.npz from the REFERENCE ITSELF (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference (read-only checkout)

Imports the reference's own eager implementations on CPU
  * src/utils/attn_ref.py:attn_ref(upcast=True) + torch.autograd          (attention fwd/bwd)
  * src/model/modeling_flash_t5.py:FlashT5LayerNorm (eager branch)        (RMSNorm)
  * src/model/modeling_flash_t5.py:FlashT5CrossEntropyLoss (torch branch) (CE + z-loss)
  * src/utils/positional_encoding.py:RelativePositionalEncoding            (T5 bias producer)
runs them on seeded inputs, asserts that the oracle/ restatements agree, and freezes inputs
and reference outputs as small fixtures.  Nothing on the GPU box reads /root/reference;
the fixtures travel instead.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("FLASHT5_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

from src.utils.attn_ref import attn_ref                                  # noqa: E402  (reference)
from src.model.modeling_flash_t5 import FlashT5LayerNorm, FlashT5CrossEntropyLoss  # noqa: E402
from src.utils.positional_encoding import RelativePositionalEncoding     # noqa: E402
from src.utils.adamw_scaled import AdamWScale                            # noqa: E402

from oracle import attn_bias_ref, rmsnorm_ref, ce_ref, adamw_ref         # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def bf16_exact(a: np.ndarray) -> torch.Tensor:
    """Round to bf16-representable fp32 so every dtype sees identical inputs."""
    return torch.from_numpy(a.astype(np.float32)).bfloat16().float()


# name: (B,H,M,N,D, bias_kind, causal, sm_scale)
ATTN_CASES = {
    "full_bias_noncausal": (2, 3, 40, 52, 16, "BH", False, 1.0),
    "bcast_batch_causal": (2, 3, 40, 52, 16, "1H", True, 1.0),
    "bcast_both_causal_m_gt_n": (1, 2, 64, 33, 32, "11", True, 0.5),
    "no_bias_default_scale": (2, 2, 48, 48, 64, None, False, None),
    "t5_bias_c1_h4": (2, 4, 128, 128, 64, "T5", False, 1.0),
    "t5_bias_causal_d128": (1, 2, 96, 96, 128, "T5", True, 1.0),
}


def make_attn_inputs(name, seed):
    B, H, M, N, D, kind, causal, scale = ATTN_CASES[name]
    rng = np.random.default_rng(seed)
    q = bf16_exact(rng.standard_normal((B, H, M, D)))
    k = bf16_exact(rng.standard_normal((B, H, N, D)))
    v = bf16_exact(rng.standard_normal((B, H, N, D)))
    do = bf16_exact(rng.standard_normal((B, H, M, D)))
    if kind is None:
        bias = None
    elif kind == "T5":
        table = bf16_exact(0.5 * rng.standard_normal((32, H)))
        pe = RelativePositionalEncoding(32, 128, H, max(M, N), bidirectional=not causal)
        with torch.no_grad():
            pe.relative_attention_bias.weight.copy_(table)
            bias = pe.compute_bias(M, N).contiguous().float()
        mine = attn_bias_ref.t5_bias(table, M, N, bidirectional=not causal)
        assert torch.equal(mine, bias), "oracle t5_bias != reference compute_bias"
    else:
        shape = {"BH": (B, H, M, N), "1H": (1, H, M, N), "11": (1, 1, M, N)}[kind]
        bias = bf16_exact(rng.standard_normal(shape))
    return q, k, v, bias, do, causal, scale


def run_reference_attn(q, k, v, bias, do, causal, scale):
    D = q.shape[-1]
    s = scale if scale is not None else 1.0 / np.sqrt(D)
    M, N = q.shape[2], k.shape[2]
    # Rows with no visible key (causal, M > N): the reference attn_ref yields NaN there (softmax of
    # all -inf) while the Triton kernel defines O = 0, L = -inf (flash_attention_v2_bias.py:470-473).
    # Those rows are the prefix i < M - N; run the reference on the valid suffix (an N x N causal
    # problem, identical maths) and zero-pad its outputs.
    r0 = max(M - N, 0) if causal else 0
    valid = torch.arange(M) >= r0
    qs = q[:, :, r0:].clone().requires_grad_(True)
    ks = k.clone().requires_grad_(True)
    vs = v.clone().requires_grad_(True)
    b = bias[:, :, r0:].clone().requires_grad_(True) if bias is not None else None
    o = attn_ref(qs, ks, vs, b, s, causal=causal, upcast=True)
    ins = [qs, ks, vs] + ([b] if b is not None else [])
    grads = torch.autograd.grad(o, ins, do[:, :, r0:])

    def pad_rows(t):
        if r0 == 0:
            return t
        z = torch.zeros(*t.shape[:2], r0, t.shape[3])
        return torch.cat([z, t], dim=2)
    dq, dk, dv = pad_rows(grads[0]), grads[1], grads[2]
    db = pad_rows(grads[3]) if b is not None else None
    return pad_rows(o.detach()), dq, dk, dv, db, valid


def gen_attn():
    for i, name in enumerate(ATTN_CASES):
        q, k, v, bias, do, causal, scale = make_attn_inputs(name, 100 + i)
        o, dq, dk, dv, db, valid = run_reference_attn(q, k, v, bias, do, causal, scale)
        # reference fp32 has NaNs in dq rows of empty rows? (they are masked out above) -> assert finite
        for t in (o, dq, dk, dv) + ((db,) if db is not None else ()):
            assert torch.isfinite(t).all(), name
        # oracle agreement (fp64 restatement vs reference fp32 eager+autograd)
        do_m = torch.where(valid.view(1, 1, -1, 1), do, torch.zeros_like(do))
        oo, LL, odq, odk, odv, odb = attn_bias_ref.attn_fwd_bwd(q, k, v, bias, do_m, causal, scale)
        for nm, a, b_ in (("o", oo, o), ("dq", odq, dq), ("dk", odk, dk), ("dv", odv, dv)):
            mx, rf = attn_bias_ref.error_metrics(a, b_)
            assert rf < 2e-6 and mx < 1e-5 * (1 + b_.abs().max().item()), (name, nm, mx, rf)
        if db is not None:
            mx, rf = attn_bias_ref.error_metrics(odb, db)
            assert rf < 2e-6 and mx < 1e-5 * (1 + db.abs().max().item()), (name, "dbias", mx, rf)
        out = dict(q=q.numpy(), k=k.numpy(), v=v.numpy(), do=do.numpy(),
                   o=o.numpy(), dq=dq.numpy(), dk=dk.numpy(), dv=dv.numpy(),
                   valid_rows=valid.numpy(), causal=np.array(causal),
                   sm_scale=np.array(np.nan if scale is None else scale, dtype=np.float64))
        if bias is not None:
            out["bias"] = bias.numpy()
            out["dbias"] = db.numpy()
        np.savez_compressed(os.path.join(GOLD, f"attn_{name}.npz"), **out)
        print(f"attn_{name}: ok  (oracle==reference)")


def gen_buckets():
    rel = torch.tensor([-300, -128, -127, -64, -32, -16, -9, -8, -7, -1, 0, 1, 7, 8, 9, 16, 32, 64, 127, 128, 300])
    bi = RelativePositionalEncoding._relative_position_bucket(rel.clone(), True, 32, 128)
    uni = RelativePositionalEncoding._relative_position_bucket(rel.clone(), False, 32, 128)
    assert torch.equal(attn_bias_ref.t5_relative_position_bucket(rel.clone(), True), bi)
    assert torch.equal(attn_bias_ref.t5_relative_position_bucket(rel.clone(), False), uni)
    np.savez_compressed(os.path.join(GOLD, "t5_buckets.npz"), rel=rel.numpy(), bidirectional=bi.numpy(),
                        unidirectional=uni.numpy())
    print("t5_buckets: ok", bi.tolist(), uni.tolist())


def gen_rmsnorm():
    rng = np.random.default_rng(7)
    for rows, n in ((12, 768), (7, 1000), (5, 512)):
        x = bf16_exact(rng.standard_normal((rows, n)) * 2.0)
        w = bf16_exact(1.0 + 0.1 * rng.standard_normal((n,)))
        dy = bf16_exact(rng.standard_normal((rows, n)))
        mod = FlashT5LayerNorm(n, eps=1e-6, use_triton_layernorm=False)
        with torch.no_grad():
            mod.weight.copy_(w)
        xr = x.clone().requires_grad_(True)
        y = mod(xr)
        dx, dw = torch.autograd.grad(y, (xr, mod.weight), dy)
        oy, rstd = rmsnorm_ref.rmsnorm_fwd(x, w, 1e-6)
        odx, odw = rmsnorm_ref.rmsnorm_bwd(dy, x, w, rstd)
        for nm, a, b_ in (("y", oy, y), ("dx", odx, dx), ("dw", odw, dw)):
            mx, rf = attn_bias_ref.error_metrics(a, b_.detach())
            assert rf < 2e-6, (rows, n, nm, mx, rf)
        np.savez_compressed(os.path.join(GOLD, f"rmsnorm_{rows}x{n}.npz"), x=x.numpy(), w=w.numpy(), dy=dy.numpy(),
                            y=y.detach().numpy(), dx=dx.numpy(), dw=dw.numpy(), eps=np.array(1e-6))
        print(f"rmsnorm_{rows}x{n}: ok")


def gen_ce():
    rng = np.random.default_rng(11)
    for rows, V, z, sm in ((9, 1000, 0.0, 0.0), (8, 4099, 1e-4, 0.0), (6, 2048, 1e-4, 0.1)):
        logits = bf16_exact(rng.standard_normal((rows, V)) * 3.0)
        labels = torch.from_numpy(rng.integers(0, V, size=(rows,))).long()
        labels[1] = -100
        mod = FlashT5CrossEntropyLoss(z_loss_factor=z, label_smoothing=sm, use_triton_crossentropy=False)
        lr = logits.clone().requires_grad_(True)
        loss = mod(lr.view(1, rows, V), labels.view(1, rows))     # mean over NON-ignored rows
        (dlogits,) = torch.autograd.grad(loss, lr)
        n_valid = int((labels != -100).sum())
        losses, zl, lse = ce_ref.ce_fwd(logits, labels, sm, 1.0, z)
        o_loss = losses.sum() / n_valid
        dl = torch.full((rows,), 1.0 / n_valid, dtype=torch.float64)
        o_dlogits = ce_ref.ce_bwd(dl, logits, lse, labels, sm, 1.0, z)
        assert abs(o_loss.item() - loss.item()) < 2e-5 * max(1.0, abs(loss.item())), (rows, V, o_loss.item(), loss.item())
        mx, rf = attn_bias_ref.error_metrics(o_dlogits, dlogits)
        assert rf < 5e-6, (rows, V, mx, rf)
        np.savez_compressed(os.path.join(GOLD, f"ce_{rows}x{V}.npz"), logits=logits.numpy(), labels=labels.numpy(),
                            loss_mean_valid=np.array(loss.item()), dlogits=dlogits.numpy(),
                            z=np.array(z), smoothing=np.array(sm))
        print(f"ce_{rows}x{V}: ok")


def gen_t5_bias():
    """compute_bias + the embedding backward of the REFERENCE module (CPU), for both bucket kinds and M != N."""
    rng = np.random.default_rng(21)
    for name, (H, M, N, bidir, nb, maxd) in {"bidir_40x56": (3, 40, 56, True, 32, 128), "unidir_64x64": (4, 64, 64, False, 32, 128),
                                             "bidir_300x300_nb16": (2, 300, 300, True, 16, 64)}.items():
        pe = RelativePositionalEncoding(nb, maxd, H, max(M, N), bidirectional=bidir)
        table = torch.from_numpy(rng.standard_normal((nb, H)).astype(np.float32))
        with torch.no_grad():
            pe.relative_attention_bias.weight.copy_(table)
        bias = pe.compute_bias(M, N)
        dbias = torch.from_numpy(rng.standard_normal((1, H, M, N)).astype(np.float32))
        (dtable,) = torch.autograd.grad(bias, pe.relative_attention_bias.weight, dbias)
        mine = attn_bias_ref.t5_bias(table, M, N, bidirectional=bidir, num_buckets=nb, max_distance=maxd)
        assert torch.equal(mine, bias.detach()), name
        np.savez_compressed(os.path.join(GOLD, f"t5bias_{name}.npz"), table=table.numpy(), bias=bias.detach().numpy(),
                            dbias=dbias.numpy(), dtable=dtable.numpy(), M=np.array(M), N=np.array(N), bidirectional=np.array(bidir),
                            num_buckets=np.array(nb), max_distance=np.array(maxd))
        print(f"t5bias_{name}: ok")


# name: (B,H,M,N,D, causal, sm_scale, num_buckets, max_distance)
RPE_CASES = {
    "enc_300x300": (2, 3, 300, 300, 32, False, 1.0, 32, 128),        # constant tiles on both sides + the band
    "dec_causal_200x200": (1, 2, 200, 200, 64, True, 1.0, 32, 128),  # unidirectional buckets, causal mask
    "cross_shape_70x330_nb16": (2, 2, 70, 330, 16, False, 0.25, 16, 64),   # M != N, other bucket geometry
}


def gen_attn_rpe():
    """The reference's dense composition of the fa2_rpe semantics: RelativePositionalEncoding.compute_bias ->
    attn_ref(upcast=True) -> autograd down to the embedding weight (the (num_buckets, H) table)."""
    for i, (name, (B, H, M, N, D, causal, scale, nb, maxd)) in enumerate(RPE_CASES.items()):
        rng = np.random.default_rng(300 + i)
        q = bf16_exact(rng.standard_normal((B, H, M, D)))
        k = bf16_exact(rng.standard_normal((B, H, N, D)))
        v = bf16_exact(rng.standard_normal((B, H, N, D)))
        do = bf16_exact(rng.standard_normal((B, H, M, D)))
        table = bf16_exact(0.5 * rng.standard_normal((nb, H)))
        pe = RelativePositionalEncoding(nb, maxd, H, max(M, N), bidirectional=not causal)
        with torch.no_grad():
            pe.relative_attention_bias.weight.copy_(table)
        qs, ks, vs = (t.clone().requires_grad_(True) for t in (q, k, v))
        bias = pe.compute_bias(M, N)
        o = attn_ref(qs, ks, vs, bias, scale, causal=causal, upcast=True)
        dq, dk, dv, dtable = torch.autograd.grad(o, [qs, ks, vs, pe.relative_attention_bias.weight], do)
        oo, LL, odq, odk, odv, odt = attn_bias_ref.attn_rpe_fwd_bwd(q, k, v, table, do, causal, scale, num_buckets=nb,
                                                                    max_distance=maxd)
        for nm, a, b_ in (("o", oo, o.detach()), ("dq", odq, dq), ("dk", odk, dk), ("dv", odv, dv), ("dtable", odt, dtable)):
            mx, rf = attn_bias_ref.error_metrics(a, b_)
            assert rf < 2e-6 and mx < 1e-5 * (1 + b_.abs().max().item()), (name, nm, mx, rf)
        np.savez_compressed(os.path.join(GOLD, f"rpe_{name}.npz"), q=q.numpy(), k=k.numpy(), v=v.numpy(), do=do.numpy(),
                            table=table.numpy(), o=o.detach().numpy(), dq=dq.numpy(), dk=dk.numpy(), dv=dv.numpy(),
                            dtable=dtable.numpy(), causal=np.array(causal), sm_scale=np.array(scale, dtype=np.float64),
                            num_buckets=np.array(nb), max_distance=np.array(maxd))
        print(f"rpe_{name}: ok  (oracle==reference)")


# name: (dtype, kahan, foreach, weight_decay, correct_bias, shapes)
ADAMW_CASES = {
    "fp32_wd": (torch.float32, False, False, 0.03, True, [(300, 7), (513,), (4, 5, 6)]),
    "fp32_foreach_nobc": (torch.float32, False, True, 0.0, False, [(300, 7), (64,)]),
    "bf16_kahan_wd": (torch.bfloat16, True, False, 0.03, True, [(300, 7), (513,), (8, 16)]),
    "bf16_plain": (torch.bfloat16, False, False, 0.0, True, [(300, 7), (33,)]),
    "bf16_kahan_foreach": (torch.bfloat16, True, True, 0.01, True, [(129, 3), (40,)]),
    "fp32_tiny_rms": (torch.float32, False, False, 0.0, True, [(50, 4)]),          # rms(p) < 1e-3: the 1e-3 floor applies
}


def gen_adamw():
    """Three steps of the REFERENCE optimizer (CPU) on seeded parameters / gradients; the oracle's dtype-faithful
    restatement must reproduce the per-tensor path bit for bit and stay within rounding of the foreach path."""
    lr, betas, eps = 2e-2, (0.9, 0.95), 1e-6
    for i, (name, (dt, kahan, foreach, wd, cb, shapes)) in enumerate(ADAMW_CASES.items()):
        rng = np.random.default_rng(500 + i)
        scale = 1e-4 if name == "fp32_tiny_rms" else 1.0
        p0 = [torch.from_numpy((scale * rng.standard_normal(s)).astype(np.float32)).to(dt) for s in shapes]
        grads = [[torch.from_numpy(rng.standard_normal(s).astype(np.float32)).to(dt) for s in shapes] for _ in range(3)]
        params = [torch.nn.Parameter(t.clone()) for t in p0]
        opt = AdamWScale(params, lr=lr, betas=betas, eps=eps, weight_decay=wd, kahan_sum=kahan, foreach=foreach, correct_bias=cb)
        # the oracle, stepped alongside
        om = [torch.zeros_like(t) for t in p0]
        ov = [torch.zeros_like(t) for t in p0]
        oc = [torch.zeros_like(t) if (kahan and dt != torch.float32) else None for t in p0]
        op = [t.clone() for t in p0]
        for step in range(3):
            for prm, g in zip(params, grads[step]):
                prm.grad = g.clone()
            opt.step()
            for j in range(len(op)):
                op[j], om[j], ov[j], oc[j] = adamw_ref.step_like_reference(op[j], grads[step][j], om[j], ov[j], oc[j], step + 1, lr,
                                                                            betas[0], betas[1], eps, wd, cb)
        out = {"dtype": np.array(str(dt)), "kahan": np.array(kahan), "foreach": np.array(foreach), "weight_decay": np.array(wd),
               "correct_bias": np.array(cb), "lr": np.array(lr), "beta1": np.array(betas[0]), "beta2": np.array(betas[1]),
               "eps": np.array(eps), "n": np.array(len(shapes))}
        for j, prm in enumerate(params):
            st = opt.state[prm]
            ref = (prm.detach(), st["exp_avg"], st["exp_avg_sq"], st["kahan_comp"])
            mine = (op[j], om[j], ov[j], oc[j])
            for nm, a, b_ in zip(("p", "m", "v", "comp"), mine, ref):
                if b_ is None:
                    assert a is None, (name, nm)
                    continue
                if not foreach:
                    assert torch.equal(a, b_), (name, j, nm, (a.float() - b_.float()).abs().max())
                else:                      # regrouped arithmetic: within a few roundings of the tensor dtype
                    tol = 2e-2 if dt == torch.bfloat16 else 1e-5
                    if nm == "comp":       # the compensation term is a rounding residual: compare the compensated value p + c
                        a, b_cmp = mine[0].float() + a.float(), ref[0].float() + b_.float()
                        tol = 2e-3
                    else:
                        b_cmp = b_.float()
                    err = (a.float() - b_cmp).abs().max().item()
                    assert err <= tol * b_cmp.abs().max().item() + 1e-12, (name, j, nm, err)
                out[f"{nm}{j}"] = b_.float().numpy()
            out[f"p0_{j}"] = p0[j].float().numpy()
            for step in range(3):
                out[f"g{step}_{j}"] = grads[step][j].float().numpy()
        np.savez_compressed(os.path.join(GOLD, f"adamw_{name}.npz"), **out)
        print(f"adamw_{name}: ok  (oracle==reference)")


if __name__ == "__main__":
    torch.manual_seed(0)
    if "--only-adamw" in sys.argv:
        gen_adamw()
        sys.exit(0)
    if "--only-rpe" in sys.argv:
        gen_attn_rpe()
        sys.exit(0)
    gen_attn_rpe()
    gen_adamw()
    gen_t5_bias()
    gen_buckets()
    gen_attn()
    gen_rmsnorm()
    gen_ce()
    print("golden fixtures written to", GOLD)
