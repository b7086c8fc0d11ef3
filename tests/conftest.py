import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure only)."""
    from oracle import cpu
    cpu.build()
    return cpu


@pytest.fixture(scope="session")
def refgpu():
    """The reference's own CUDA kernels (oracle/_ref), if they were built in the authoring
    container; GPU tests that need them skip otherwise."""
    from tests import refgpu as r
    if not r.available():
        pytest.skip("oracle/_ref not built (needs /root/reference: make -C oracle ref)")
    return r
