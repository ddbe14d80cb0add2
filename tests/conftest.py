"""pytest configuration: `gpu` marker (tests that need a B200) and shared scene/oracle helpers.

CPU tests (`-m "not gpu"`) cover the oracle against the reference's analytic assertions and the committed golden
vectors, the host-side model compiler, and that libgsrigid.so loads and exports every symbol of include/gsrigid.h.
GPU tests (`-m gpu`) are the parity tests proper: CUDA path (through the C ABI) vs the oracle on the same seeded
inputs.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may touch oracle/.
"""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# the reference's own single-precision tolerance (tests/conftest.py:80 TOL_SINGLE)
TOL_SINGLE = 5e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def gs():
    import genesis_b200 as gs

    gs.init(backend=gs.gpu, seed=0)
    return gs


@pytest.fixture(scope="session", autouse=True)
def _built_libs():
    """Build both shared libraries once (nvcc cross-compiles without a GPU; the oracle is plain g++)."""
    from genesis_b200 import _lib
    from oracle import pyoracle

    _lib.build()
    pyoracle.build()
