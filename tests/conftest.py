import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def tvk():
    """A CUDA context through the C ABI; fails loudly when the library or the GPU is missing."""
    from imagevis3d_b200.context import TvkContext
    ctx = TvkContext(0)
    yield ctx
    ctx.close()
