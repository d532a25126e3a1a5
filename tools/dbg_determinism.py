import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flasht5_b200  # noqa
dev = "cuda:0"
def run(B, H, S, D, bias_on, causal=False, reps=3, mode2=False):
    g = torch.Generator(device=dev).manual_seed(2)
    mk = lambda: torch.randn(B, S, H, D, generator=g, device=dev).to(torch.bfloat16).permute(0, 2, 1, 3)
    q, k, v, do = mk(), mk(), mk(), mk()
    bias = (0.5 * torch.randn(1, H, S, S + (4 if mode2 else 0), generator=g, device=dev)).to(torch.bfloat16)[..., :S] if bias_on else None
    o0, L0 = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, causal, 1.0)
    for r in range(reps):
        o, L = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, causal, 1.0)
        neq = (o != o0)
        nL = (L != L0)
        if neq.any() or nL.any():
            idx = neq.any(-1).nonzero()
            print(f"  rep {r}: o mismatches {int(neq.sum())} elems in {idx.shape[0]} rows; L mismatches {int(nL.sum())}; first rows {idx[:6].tolist()}; maxdiff {float((o.float()-o0.float()).abs().max())}")
            rows = idx[:, 2] if idx.numel() else idx
            if idx.numel():
                print("   row%128 histogram (first 16):", torch.bincount(rows % 128, minlength=128)[:16].tolist(), " mblocks:", torch.bincount(rows // 128).tolist())
        else:
            print(f"  rep {r}: identical")
    # single-batch vs full
    o5, L5 = torch.ops.b200t5.attn_bias_fwd(q[1:2], k[1:2], v[1:2], bias, causal, 1.0)
    print("  slice-vs-full o equal:", bool(torch.equal(o5, o0[1:2])), " L equal:", bool(torch.equal(L5, L0[1:2])))
print("env", {k: v for k, v in os.environ.items() if k.startswith("B200T5")})
print("cfg headline mode1"); run(32, 8, 1024, 64, True)
print("cfg headline mode2"); run(32, 8, 1024, 64, True, mode2=True)
