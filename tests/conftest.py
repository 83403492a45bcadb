import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))  # orclib: bindings of the test-only oracle
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "ref_solves.npz"))


@pytest.fixture(scope="session")
def plan_unordered():
    return np.load(os.path.join(GOLDEN, "plan_unordered.npz"))


@pytest.fixture(scope="session")
def plan_reordered():
    return np.load(os.path.join(GOLDEN, "plan_reordered.npz"))
