import os
import sys

import pytest

# The parity suite pins the BIT-EXACT kernels against the oracle: mid-n float runs (13 <= n <= 55) take the
# warp-per-problem kernel here (also in the C++ test binaries, which inherit the environment).  The library default for
# those sizes — the tensor-core kernel of wtc.cuh, tolerance-held — is pinned by tests/test_gpu_wtc.py, which switches it
# on explicitly (Context.set_exact(False)).
os.environ.setdefault("TOB200_WPP_TC", "0")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible, so a bare `pytest tests/`
    works on the CPU container; `-m gpu` on the box runs them for real."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
