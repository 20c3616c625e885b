import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """The in-tree native artefacts (libyota_cuda.so, liboracle.so)."""
    import __graft_entry__ as g

    from yota_b200 import ffi

    if not os.path.exists(ffi.LIB_PATH):
        g.build()
    return ffi.lib()
