"""GPU-box tool: four forward + backward calls of the headline attention shape (B=32 H=8 S=1024 d=64 bf16) for ncu captures.
    ncu --set full -k regex:attn_bwd_kernel_v3 -s 2 -c 1 -o gpurun_out/x python tools/headline_step.py [bias|nobias]"""
import sys, torch
sys.path.insert(0, ".")
import flasht5_b200
mode = sys.argv[1] if len(sys.argv) > 1 else "bias"
DEV = "cuda:0"
B, H, S, D = 32, 8, 1024, 64
g = torch.Generator(device=DEV).manual_seed(1)
mk = lambda: torch.randn(B, S, H, D, generator=g, device=DEV).to(torch.bfloat16).permute(0, 2, 1, 3)
q, k, v, do = mk(), mk(), mk(), mk()
bias = torch.randn(1, H, S, S, generator=g, device=DEV).to(torch.bfloat16) if mode == "bias" else None
for i in range(4):
    o, L = torch.ops.b200t5.attn_bias_fwd(q, k, v, bias, False, 1.0)
    torch.ops.b200t5.attn_bias_bwd(o, do, q, k, v, bias, L, False, 1.0)
torch.cuda.synchronize()
