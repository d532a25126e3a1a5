"""Oracle for the RMSNorm op (TEST INFRASTRUCTURE, not product).

Restates /root/reference/src/model/ops/rms_norm.py:
  forward  :25-66   rstd = 1/sqrt(mean(x^2) + eps) (fp32 per row); y = x * rstd * w
  backward :68-131  xhat = x*rstd; wdy = w*dy; c1 = mean(xhat*wdy); dx = (wdy - xhat*c1)*rstd;
                    dw = sum_rows dy*xhat (fp32 accumulation, cast to w dtype at :234)
Pinned against the reference's eager module math (modeling_flash_t5.py:105-112) by
oracle/make_golden.py -> tests/golden/rmsnorm_*.npz.
"""
import torch


def rmsnorm_fwd(x, w, eps=1e-6, dtype=torch.float64):
    xf, wf = x.to(dtype), w.to(dtype)
    rstd = 1.0 / torch.sqrt((xf * xf).mean(-1, keepdim=True) + eps)
    return xf * rstd * wf, rstd.squeeze(-1)


def rmsnorm_bwd(dy, x, w, rstd, dtype=torch.float64):
    xf, wf, dyf = x.to(dtype), w.to(dtype), dy.to(dtype)
    r = rstd.to(dtype).unsqueeze(-1)
    xhat = xf * r
    wdy = wf * dyf
    c1 = (xhat * wdy).mean(-1, keepdim=True)
    dx = (wdy - xhat * c1) * r
    dw = (dyf * xhat).reshape(-1, x.shape[-1]).sum(0)
    return dx, dw
