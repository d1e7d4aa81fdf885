import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_heatmap():
    return np.load(os.path.join(GOLDEN, "heatmap.npz"))


@pytest.fixture(scope="session")
def golden_gru():
    return np.load(os.path.join(GOLDEN, "gru.npz"))


@pytest.fixture(scope="session")
def real_traces():
    return np.load(os.path.join(GOLDEN, "real_traces.npz"))


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree library, (re)built if stale.  nvcc cross-compiles without a GPU."""
    from roomslam_b200 import build
    return build.build_library()
