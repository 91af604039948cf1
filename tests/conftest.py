import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_points():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_points.npz"))


@pytest.fixture(scope="session")
def golden_calls():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_calls.npz"))


@pytest.fixture(scope="session")
def built():
    """Build (if stale) the CUDA library and the test-only host emulation."""
    import __graft_entry__ as ge
    ge.build()
    return True
