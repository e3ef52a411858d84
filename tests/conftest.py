import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100a device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree C-ABI library, (re)built if sources are newer (nvcc cross-compiles without a GPU)."""
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import _lib
    _lib.build()
    return _lib.load()


def golden(name):
    import numpy as np
    return np.load(os.path.join(REPO, "tests", "golden", name + ".npz"))
