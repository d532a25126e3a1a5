import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real sm_100 (B200) GPU; run with `-m gpu` on the GPU box")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly (never skip silently) when selected with -m gpu on a box without a GPU;
    when not explicitly selected on a CPU-only box they are skipped."""
    import torch
    if torch.cuda.is_available():
        return
    selected_gpu = "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or "")
    if selected_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device (GPU tests run on the B200 box with -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (built on demand in the CPU container; prebuilt on the GPU box)."""
    from flasht5_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        from flasht5_b200.build import build
        build()
    return _cabi.load()
