"""TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference hot path (catie-aq/flashT5 attention-with-bias,
RMSNorm, cross-entropy + z-loss).  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this package; the
product path (`flasht5_b200/`) never does and has no CPU fallback.
"""
