"""GPU-box tool: the persistent backward under repetition -- many work items per CTA, short causal items, L2 flushed every third
call, dK / dV compared bitwise every 50 calls.  (Found: a second arrival on an open mbarrier phase with one-tile items.)
    python tools/bwd_stress.py [iterations]"""
import sys, os, torch
sys.path.insert(0, ".")
import flasht5_b200
DEV = "cuda:0"
ops = torch.ops.b200t5
flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device=DEV)
def run(B, H, S, D, bias_on, causal, iters):
    g = torch.Generator(device=DEV).manual_seed(1)
    mk = lambda: torch.randn(B, S, H, D, generator=g, device=DEV).to(torch.bfloat16).permute(0, 2, 1, 3)
    q, k, v, do = mk(), mk(), mk(), mk()
    bias = torch.randn(1, H, S, S, generator=g, device=DEV).to(torch.bfloat16) if bias_on else None
    o, L = ops.attn_bias_fwd(q, k, v, bias, causal, 1.0)
    ref = None
    for i in range(iters):
        if i % 3 == 0: flush.zero_()
        out = ops.attn_bias_bwd(o, do, q, k, v, bias, L, causal, 1.0)
        if i % 50 == 0:
            torch.cuda.synchronize()
            if ref is None: ref = [t.clone() for t in out[1:3]]
            else: assert all(torch.equal(a, b) for a, b in zip(ref, out[1:3])), "dk/dv changed"
    torch.cuda.synchronize()
    print("ok", B, H, S, D, bias_on, causal, iters, flush=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
run(16, 12, 1024, 64, False, False, n)
run(16, 12, 1024, 64, True, True, n)
run(16, 12, 1024, 64, False, True, n)
run(32, 8, 1024, 64, True, False, n)
run(8, 16, 2048, 64, True, True, n // 3)
run(4, 4, 640, 32, True, True, n)
