import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# The oracle's torch graphs are thousands of tiny ops (10^3 .. 10^4 points, width 20): with one thread per core the intra-op
# pool spends its time spinning (measured here: a reference replay 114 s with 8 threads, 12 s with <= 4).  Must be set
# before torch starts its pools; spawned ranks of the gloo tests inherit it.
os.environ.setdefault("OMP_NUM_THREADS", str(min(4, os.cpu_count() or 1)))
os.environ.setdefault("MKL_NUM_THREADS", os.environ["OMP_NUM_THREADS"])
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
