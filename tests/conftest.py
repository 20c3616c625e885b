import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def gpu_count():
    """Number of CUDA devices the driver offers (0 without a driver): asked of libcuda directly so
    that collection does not import torch."""
    import ctypes

    try:
        cu = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if cu.cuInit(0) != 0 or cu.cuDeviceGetCount(ctypes.byref(n)) != 0:
            return 0
        return n.value
    except OSError:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a GPU skips the `gpu` tests instead of failing them."""
    if gpu_count() > 0:
        return
    skip = pytest.mark.skip(reason="needs a B200 (no CUDA device on this host)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """The in-tree native artefacts (libyota_cuda.so, liboracle.so)."""
    import __graft_entry__ as g

    from yota_b200 import ffi

    if not os.path.exists(ffi.LIB_PATH):
        g.build()
    return ffi.lib()
