"""GPU-box tool: time the headline forward + backward step (one CUDA graph per input set, replayed) for the library named by
B200T5_LIB (default: the in-tree build).  For A/B runs inside ONE gpurun call (fresh boxes differ by +-2 %):
    for l in base new base new; do B200T5_LIB=$PWD/flasht5_b200/libb200t5_$l.so python tools/ab_step_time.py; done"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flasht5_b200  # noqa: E402,F401

DEV = "cuda:0"
B, H, S, D = 32, 8, 1024, 64
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = torch.Generator(device=DEV).manual_seed(1)
mk = lambda: torch.randn(B, S, H, D, generator=g, device=DEV).to(torch.bfloat16).permute(0, 2, 1, 3)   # noqa: E731
sets = [(mk(), mk(), mk(), (0.5 * torch.randn(1, H, S, S, generator=g, device=DEV)).to(torch.bfloat16), mk()) for _ in range(3)]
ops = torch.ops.b200t5


def kernels(i, bias_on=True):
    q, k, v, bias, do = sets[i % 3]
    bias = bias if bias_on else None
    o, L = ops.attn_bias_fwd(q, k, v, bias, False, 1.0)
    return ops.attn_bias_bwd(o, do, q, k, v, bias, L, False, 1.0)


out = []
for bias_on in (True, False):
    for i in range(3):
        kernels(i, bias_on)
    torch.cuda.synchronize()
    graphs = []
    for i in range(3):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            keep = kernels(i, bias_on)
        graphs.append((gr, keep))
    for gr, _ in graphs:
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        graphs[i % 3][0].replay()
    e1.record()
    torch.cuda.synchronize()
    out.append("%s %.1f us/step" % ("bias" if bias_on else "nobias", 1e3 * e0.elapsed_time(e1) / steps))
print(os.path.basename(os.environ.get("B200T5_LIB", "in-tree")), " | ".join(out), flush=True)
