import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "convolver_golden.npz"))


@pytest.fixture(scope="session", autouse=True)
def _oracle_built():
    """The CPU checker is compiled on demand (gcc, seconds)."""
    from oracle import bindings
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        bindings.build()
