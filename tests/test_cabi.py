"""CPU: the C-ABI library builds, loads and exports every symbol include/b200t5.h declares;
entry points that need no GPU behave; compute entry points refuse loudly without one."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT
from flasht5_b200 import _cabi


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200t5.h")).read()
    return sorted(set(re.findall(r"B200T5_API\s+[\w\s\*]+?\b(b200t5_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = _declared_symbols()
    for must in ("b200t5_attn_fwd", "b200t5_attn_bwd", "b200t5_attn_bwd_workspace_bytes", "b200t5_rmsnorm_fwd",
                 "b200t5_rmsnorm_bwd", "b200t5_ce_fwd", "b200t5_ce_bwd", "b200t5_last_error", "b200t5_abi_version"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    raw = C.CDLL(_cabi.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(raw, name), f"{name} declared in include/b200t5.h but not exported"
    assert sorted(_cabi.SYMBOLS) == _declared_symbols(), "ctypes binding and header disagree"


def test_abi_version_and_struct_layout(lib):
    assert lib.b200t5_abi_version() == _cabi.ABI_VERSION
    # layout of b200t5_attn_params as the C compiler sees it: 12 int32/float + pointer-aligned tail
    assert _cabi.AttnParams.stream.offset == 48
    assert _cabi.AttnParams.q.offset == 56
    assert _cabi.AttnParams.q_strides.offset == 64
    assert C.sizeof(_cabi.AttnParams) == 56 + 5 * 40 + 8 + 5 * 40 + 16


def test_struct_layout_matches_c_compiler(tmp_path):
    """Compile a tiny C program against the header and compare offsets with the ctypes mirror."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "off.c"
    fields = ["B", "sm_scale", "stream", "q", "q_strides", "bias", "o", "lse", "dout", "dq", "dbias", "workspace",
              "workspace_bytes"]
    body = "".join(f'printf("%zu\\n", offsetof(b200t5_attn_params, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b200t5.h"\nint main(){' + body +
                   'printf("%zu\\n", sizeof(b200t5_attn_params));return 0;}')
    exe = tmp_path / "off"
    subprocess.check_call([cc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    mine = [getattr(_cabi.AttnParams, f).offset for f in fields] + [C.sizeof(_cabi.AttnParams)]
    assert out == mine


def test_workspace_size_queries_need_no_gpu(lib):
    p = _cabi.AttnParams()
    p.B, p.H, p.M, p.N, p.D = 2, 3, 100, 77, 64
    p.bias = None
    no_bias = lib.b200t5_attn_bwd_workspace_bytes(C.byref(p))
    rows = 2 * 3 * 100
    assert no_bias >= rows * 4 + rows * 64 * 2      # delta (fp32) + one 16-bit dQ group
    p.bias = 16                                     # any non-NULL pointer: size only depends on presence
    with_bias = lib.b200t5_attn_bwd_workspace_bytes(C.byref(p))
    assert with_bias >= no_bias + rows * 80 * 2      # dS rows padded to a multiple of 8 columns
    assert lib.b200t5_rmsnorm_bwd_workspace_bytes(768) >= 148 * 768 * 4
    assert lib.b200t5_attn_bwd_workspace_bytes(None) == 0


def test_argument_validation_precedes_any_cuda_work(lib):
    p = _cabi.AttnParams()
    p.B, p.H, p.M, p.N, p.D = 1, 1, 8, 8, 48        # unsupported head dim (reference asserts the same, :233-234)
    p.dtype = _cabi.BF16
    assert lib.b200t5_attn_fwd(C.byref(p)) == -2
    assert "head dim 48" in _cabi.last_error()
    p.D = 64
    assert lib.b200t5_attn_fwd(C.byref(p)) == -1     # NULL operand pointers
    assert "non-NULL" in _cabi.last_error()
    assert lib.b200t5_attn_fwd(None) == -1
    assert lib.b200t5_rmsnorm_fwd(None, None, None, None, 1, 8, 8, 8, 1e-6, 1, 1, 0, None) == -1
    assert lib.b200t5_ce_fwd(None, None, None, None, None, 0, 1, 8, 8, 0.0, 1.0, 0.0, -100, 1, 0, None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback_anywhere():
    """Without a GPU the ops must raise, never compute."""
    import flasht5_b200 as ft
    q = torch.randn(1, 2, 16, 64, dtype=torch.bfloat16)
    with pytest.raises((RuntimeError, NotImplementedError)):
        ft.flash_attention_v2_bias(q, q, q, None)
    with pytest.raises((RuntimeError, NotImplementedError)):
        ft.fast_rms_layernorm(torch.randn(4, 64), torch.ones(64), 1e-6)
    with pytest.raises((RuntimeError, NotImplementedError)):
        ft.cross_entropy_loss(torch.randn(4, 64), torch.zeros(4, dtype=torch.long))
    # product code never imports the oracle
    import subprocess
    import sys
    code = "import sys, flasht5_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
