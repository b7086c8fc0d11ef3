import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda():
    import torch
    return torch.cuda.is_available()


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure only)."""
    from oracle import cpu
    cpu.build()
    return cpu


@pytest.fixture(scope="session")
def refgpu():
    """The reference's own CUDA kernels (oracle/_ref, built by `make -C oracle ref` where
    /root/reference exists; the .so files travel to the GPU box with the snapshot). On a box WITH a
    GPU their absence is an error, not a skip: every bitwise-against-the-reference-kernel test
    would otherwise vanish with a green exit."""
    from tests import refgpu as r
    if not r.available():
        if _cuda():
            pytest.fail("oracle/_ref/libref_*.so missing on a CUDA box: build them in the authoring "
                        "container (make -C oracle ref) so that they travel with the snapshot")
        pytest.skip("oracle/_ref not built (needs /root/reference: make -C oracle ref)")
    return r


@pytest.fixture(scope="session")
def ref_root():
    """A MoCoPCI checkout to import the reference's own Python from: /root/reference in the
    authoring container, else the copy baseline/fetch_ref.py left under baseline/_ref."""
    from baseline import fetch_ref
    root = fetch_ref.root()
    if root is None:
        if _cuda():
            pytest.fail("no reference checkout (baseline/_ref missing on a CUDA box: run "
                        "python baseline/fetch_ref.py in the authoring container)")
        pytest.skip("no reference checkout available")
    return root
