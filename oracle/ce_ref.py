"""Oracle for cross-entropy + z-loss (TEST INFRASTRUCTURE, not product).

Restates /root/reference/src/model/ops/cross_entropy_loss.py (single-rank path,
process_group=None; the vocab-parallel branch :324-351 is dead code in the reference):
  forward  :40-111   x = logits*logit_scale (fp32); lse = ln sum exp x;
                     ignore_index -> loss = z = 0;
                     loss = lse - x[label]                                    (no smoothing)
                     loss = lse - eps*sum(x)/V - (1-eps)*x[label]             (smoothing eps, :90-95)
                     z = lse_square_scale * lse^2; loss += z                  (:105-106)
  backward :119-162  dlogits = dloss*logit_scale*( softmax(x)*(1 + 2*lambda*lse)
                               - onehot(label)*(1-eps) - eps/V ),  0 on ignored rows
Pinned against torch.nn.functional.cross_entropy + the reference's compute_zloss
(modeling_flash_t5.py:55-59) by oracle/make_golden.py -> tests/golden/ce_*.npz.
"""
import torch


def ce_fwd(logits, labels, smoothing=0.0, logit_scale=1.0, lse_square_scale=0.0,
           ignore_index=-100, dtype=torch.float64):
    x = logits.to(dtype) * logit_scale
    n_rows, V = x.shape
    lse = torch.logsumexp(x, dim=-1)
    ignored = labels == ignore_index
    safe = torch.where(ignored, torch.zeros_like(labels), labels)
    x_label = x.gather(1, safe.view(-1, 1)).squeeze(1)
    if smoothing > 0.0:
        loss = lse - smoothing * x.sum(-1) / V - (1.0 - smoothing) * x_label
    else:
        loss = lse - x_label
    z = lse_square_scale * lse * lse
    loss = loss + z
    zero = torch.zeros_like(loss)
    return torch.where(ignored, zero, loss), torch.where(ignored, zero, z), lse


def ce_bwd(dloss, logits, lse, labels, smoothing=0.0, logit_scale=1.0, lse_square_scale=0.0,
           ignore_index=-100, dtype=torch.float64):
    x = logits.to(dtype) * logit_scale
    n_rows, V = x.shape
    lse = lse.to(dtype)
    probs = torch.exp(x - lse.unsqueeze(-1))
    probs = probs + 2.0 * lse_square_scale * lse.unsqueeze(-1) * probs
    ignored = labels == ignore_index
    safe = torch.where(ignored, torch.zeros_like(labels), labels)
    onehot = torch.zeros_like(probs).scatter_(1, safe.view(-1, 1), 1.0)
    if smoothing > 0.0:
        probs = probs - onehot * (1.0 - smoothing) - smoothing / V
    else:
        probs = probs - onehot
    dl = torch.where(ignored, torch.zeros_like(lse), dloss.to(dtype))
    return (dl * logit_scale).unsqueeze(-1) * probs
