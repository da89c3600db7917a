import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # A `gpu` test on a box without CUDA is skipped (the driver's CPU pass uses -m "not gpu" anyway).
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionstart(session):
    """The product library is built in-tree by `__graft_entry__.build()` (git-ignored, so a fresh checkout has none).
    Build it here when it is missing so the suite can run on its own; the engine itself never builds or falls back."""
    lib = os.path.join(ROOT, "chrono-mind_b200", "libchronomind_b200.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "chrono-mind_b200", "csrc"), "-j8", "-s"])
