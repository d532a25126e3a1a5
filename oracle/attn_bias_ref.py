"""Oracle for FlashAttention-with-additive-bias (TEST INFRASTRUCTURE, not product).

A plain CPU restatement (torch CPU tensors used as an ndarray library, explicit
formulas, no autograd) of what the reference operator computes.

Follows (paths relative to /root/reference):
  * forward semantics        src/utils/attn_ref.py:3-29  (softmax(QK^T*s + bias [+causal]) V)
  * LSE definition           src/model/ops/flash_attention_v2_bias.py:476   (L = m + ln l, natural log)
  * causal alignment         src/model/ops/flash_attention_v2_bias.py:447-449 (bottom-right, P_SEQ = N - M)
  * empty rows               src/model/ops/flash_attention_v2_bias.py:470-473 (O = 0, L = -inf)
  * bias broadcast           src/model/ops/flash_attention_v2_bias.py:46-52  (size-1 batch/head dims)
  * backward                 src/model/ops/flash_attention_v2_bias.py:516-905:
        delta = rowsum(O * dO)        :550
        P  = exp(S - L)               :690
        dV = P^T dO                   :702
        dP = dO V^T                   :710
        dS = P * (dP - delta)         :713-720
        dK = s * dS^T Q               :722,739
        dQ = s * dS K                 :893,901
        dBias = dS summed over every broadcast bias dim (:214-215 does batch only;
                the head-broadcast case is a race in the reference, SURVEY.md section 4 --
                the oracle implements the mathematically correct sum).

Parity pin: `oracle/make_golden.py` checks this file against the reference's own
`attn_ref` + torch autograd (imported from /root/reference in the build container)
and freezes the outputs in tests/golden/attn_*.npz.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch


def _expand_bias(bias: Optional[torch.Tensor], B: int, H: int, M: int, N: int, dtype) -> Optional[torch.Tensor]:
    if bias is None:
        return None
    assert bias.dim() == 4 and bias.shape[2] == M and bias.shape[3] == N
    assert bias.shape[0] in (1, B) and bias.shape[1] in (1, H)
    return bias.to(dtype).expand(B, H, M, N)


def attn_fwd(q, k, v, bias, causal: bool = False, sm_scale: Optional[float] = None,
             dtype=torch.float64) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (O, L) with O:(B,H,M,D) and L:(B,H,M) = logsumexp of the masked scores."""
    B, H, M, D = q.shape
    N = k.shape[2]
    if sm_scale is None:
        sm_scale = 1.0 / math.sqrt(D)
    qf, kf, vf = q.to(dtype), k.to(dtype), v.to(dtype)
    s = torch.matmul(qf, kf.transpose(2, 3)) * sm_scale
    bf = _expand_bias(bias, B, H, M, N, dtype)
    if bf is not None:
        s = s + bf
    if causal:
        ms = torch.arange(M).unsqueeze(-1)
        ns = torch.arange(N)
        s = torch.where(ms + (N - M) >= ns, s, torch.full_like(s, float("-inf")))
    L = torch.logsumexp(s, dim=-1)                       # -inf on rows with no visible key
    p = torch.exp(s - torch.where(torch.isinf(L), torch.zeros_like(L), L).unsqueeze(-1))
    p = torch.where(torch.isinf(L).unsqueeze(-1), torch.zeros_like(p), p)
    o = torch.matmul(p, vf)
    return o, L


def attn_bwd(q, k, v, bias, o, L, do, causal: bool = False, sm_scale: Optional[float] = None,
             dtype=torch.float64):
    """Returns (dQ, dK, dV, dBias|None) following the reference backward formulas."""
    B, H, M, D = q.shape
    N = k.shape[2]
    if sm_scale is None:
        sm_scale = 1.0 / math.sqrt(D)
    qf, kf, vf, of, dof = (t.to(dtype) for t in (q, k, v, o, do))
    Lf = L.to(dtype)
    s = torch.matmul(qf, kf.transpose(2, 3)) * sm_scale
    bf = _expand_bias(bias, B, H, M, N, dtype)
    if bf is not None:
        s = s + bf
    if causal:
        ms = torch.arange(M).unsqueeze(-1)
        ns = torch.arange(N)
        visible = (ms + (N - M) >= ns)
    else:
        visible = torch.ones(M, N, dtype=torch.bool)
    empty = torch.isinf(Lf)
    p = torch.exp(s - torch.where(empty, torch.zeros_like(Lf), Lf).unsqueeze(-1))
    p = torch.where(visible & ~empty.unsqueeze(-1), p, torch.zeros_like(p))
    delta = (of * dof).sum(-1, keepdim=True)
    dv = torch.matmul(p.transpose(2, 3), dof)
    dp = torch.matmul(dof, vf.transpose(2, 3))
    ds = p * (dp - delta)
    dk = torch.matmul(ds.transpose(2, 3), qf) * sm_scale
    dq = torch.matmul(ds, kf) * sm_scale
    dbias = None
    if bias is not None:
        dbias = ds
        if bias.shape[0] == 1 and B > 1:
            dbias = dbias.sum(0, keepdim=True)
        if bias.shape[1] == 1 and H > 1:
            dbias = dbias.sum(1, keepdim=True)
    return dq, dk, dv, dbias


def attn_fwd_bwd(q, k, v, bias, do, causal=False, sm_scale=None, dtype=torch.float64):
    o, L = attn_fwd(q, k, v, bias, causal, sm_scale, dtype)
    dq, dk, dv, db = attn_bwd(q, k, v, bias, o, L, do, causal, sm_scale, dtype)
    return o, L, dq, dk, dv, db


# ---------------------------------------------------------------------------------------------
# Eager low-precision path = the reference's `attn_ref(upcast=False)` semantics, used for
# (a) the reference's own tolerance rule  err_new <= 2*err_eager + 1e-5
#     (tests/fa2_triton/test_fa2_bias.py:28,64-67) and
# (b) the cpu_baseline / `--impl reference` leg of bench.py (torch autograd, all host threads).
# Restates src/utils/attn_ref.py:3-29 op for op (matmul in input dtype, softmax in fp32, cast back).
# ---------------------------------------------------------------------------------------------
def attn_eager_lowp(q, k, v, bias, causal=False, sm_scale=None):
    B, H, M, D = q.shape
    N = k.shape[2]
    if sm_scale is None:
        sm_scale = 1.0 / math.sqrt(D)
    p = torch.matmul(q, k.transpose(2, 3))
    p = p * sm_scale
    if bias is not None:
        p = p + bias
    if causal:
        ms = torch.arange(M, device=q.device).unsqueeze(-1)
        ns = torch.arange(N, device=q.device)
        p = torch.where(ms + N - M >= ns, p, float("-inf"))
    p = torch.softmax(p.float(), dim=-1).to(q.dtype)
    return torch.matmul(p, v)


def error_metrics(x: torch.Tensor, ref: torch.Tensor):
    """max|delta| and relative Frobenius norm, the two figures BASELINE.json asks for."""
    x = x.detach().double().cpu()
    ref = ref.detach().double().cpu()
    finite = torch.isfinite(ref)
    if not torch.equal(torch.isfinite(x), finite):
        return float("inf"), float("inf")
    d = torch.where(finite, x - ref, torch.zeros_like(ref))
    r = torch.where(finite, ref, torch.zeros_like(ref))
    maxabs = d.abs().max().item() if d.numel() else 0.0
    denom = r.norm().item()
    relf = d.norm().item() / denom if denom > 0 else d.norm().item()
    return maxabs, relf


# ---------------------------------------------------------------------------------------------
# T5 relative-position bias producer (the hot path's `bias` input), restating
# src/utils/positional_encoding.py:25-71 (`_relative_position_bucket`) and :73-102 (compute_bias).
# Known-answer vector recorded in SURVEY.md section 8c and checked in tests/test_oracle.py.
# ---------------------------------------------------------------------------------------------
def t5_relative_position_bucket(relative_position: torch.Tensor, bidirectional=True,
                                num_buckets=32, max_distance=128) -> torch.Tensor:
    relative_buckets = torch.zeros_like(relative_position)
    if bidirectional:
        num_buckets //= 2
        relative_buckets = relative_buckets + (relative_position > 0).long() * num_buckets
        relative_position = relative_position.abs()
    else:
        relative_position = -torch.min(relative_position, torch.zeros_like(relative_position))
    max_exact = num_buckets // 2
    is_small = relative_position < max_exact
    rp_large = max_exact + (
        torch.log(relative_position.float().clamp(min=1) / max_exact)
        / math.log(max_distance / max_exact) * (num_buckets - max_exact)
    ).long()
    rp_large = torch.min(rp_large, torch.full_like(rp_large, num_buckets - 1))
    return relative_buckets + torch.where(is_small, relative_position, rp_large)


def t5_bias(table: torch.Tensor, M: int, N: int, bidirectional=True, num_buckets=32,
            max_distance=128) -> torch.Tensor:
    """table: (num_buckets, H) -> dense bias (1, H, M, N) (Toeplitz in j - i)."""
    ctx = torch.arange(M).unsqueeze(-1)
    mem = torch.arange(N).unsqueeze(0)
    buckets = t5_relative_position_bucket(mem - ctx, bidirectional, num_buckets, max_distance)
    vals = table[buckets]                    # (M, N, H)
    return vals.permute(2, 0, 1).unsqueeze(0).contiguous()


# ---------------------------------------------------------------------------------------------
# Attention with the T5 bias taken straight from the embedding table (the reference's "fa2_rpe" surface,
# src/model/modeling_flash_t5.py:275-279).  The flash-attention fork behind that call is not in the reference
# checkout, so the oracle restates the reference's own dense composition: compute_bias (:73-102 of
# src/utils/positional_encoding.py) -> cast to the attention dtype -> attention -> and, backwards, the
# embedding's scatter-add of dBias into the (num_buckets, H) table.
# ---------------------------------------------------------------------------------------------
def t5_dtable(dbias: torch.Tensor, M: int, N: int, bidirectional=True, num_buckets=32, max_distance=128) -> torch.Tensor:
    """dbias (1, H, M, N) -> dtable (num_buckets, H): sum of dbias over the positions of each bucket."""
    ctx = torch.arange(M).unsqueeze(-1)
    mem = torch.arange(N).unsqueeze(0)
    buckets = t5_relative_position_bucket(mem - ctx, bidirectional, num_buckets, max_distance)   # (M, N)
    H = dbias.shape[1]
    dtable = torch.zeros(num_buckets, H, dtype=dbias.dtype)
    dtable.index_add_(0, buckets.reshape(-1), dbias[0].permute(1, 2, 0).reshape(M * N, H))
    return dtable


def attn_rpe_fwd_bwd(q, k, v, table, do, causal=False, sm_scale=None, bidirectional=None, num_buckets=None,
                     max_distance=128, bias_dtype=None, dtype=torch.float64):
    """-> (o, L, dq, dk, dv, dtable).  table: (num_buckets, H).  bias_dtype: round the gathered bias to this dtype
    first (what the dense path does when it casts the bias to the q dtype); None keeps it exact."""
    M, N = q.shape[2], k.shape[2]
    if bidirectional is None:
        bidirectional = not causal
    if num_buckets is None:
        num_buckets = table.shape[0]
    bias = t5_bias(table, M, N, bidirectional, num_buckets, max_distance)
    if bias_dtype is not None:
        bias = bias.to(bias_dtype)
    o, L, dq, dk, dv, dbias = attn_fwd_bwd(q, k, v, bias.to(dtype), do, causal, sm_scale, dtype=dtype)
    return o, L, dq, dk, dv, t5_dtable(dbias, M, N, bidirectional, num_buckets, max_distance)
