import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """The CUDA extension is git-ignored (it travels with the working tree): build it in-tree if a fresh checkout is
    being tested.  This only compiles; nothing here provides a substitute for the kernels."""
    from smallhardface_b200 import lib
    try:
        if not os.path.exists(lib.LIB_PATH):
            raise lib.ShfError("missing")
        lib.load()                                   # raises on a stale ABI
    except lib.ShfError:
        lib.build()
    return lib.LIB_PATH

