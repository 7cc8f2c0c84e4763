import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def smm_lib():
    """The C-ABI shared library, built in-tree if stale (nvcc cross-compiles without a GPU)."""
    from smmregrid_b200 import _build, _lib
    _build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure): numpy restatement + C port of the reference."""
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
