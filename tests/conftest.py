import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """The tests load the in-tree libraries (product C ABI, host simulator, compiled reference).  They are built by
    __graft_entry__.build(); do that here if a fresh checkout has none yet (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    lib = os.path.join(ROOT, "hmp3_b200", "_lib")
    need = [os.path.join(lib, "libhmp3_b200.so"), os.path.join(lib, "libhmp3_sim.so")]
    if os.path.isdir(ge.REF_SRC):
        need.append(os.path.join(ROOT, "oracle", "_ref", "libhmp3ref.so"))
    if not all(os.path.exists(p) for p in need):
        ge.build()
