import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_tables():
    return np.load(os.path.join(GOLDEN, "tables.npz"))


@pytest.fixture(scope="session")
def golden_get_cost():
    return np.load(os.path.join(GOLDEN, "get_cost.npz"))


@pytest.fixture(scope="session")
def golden_lethal():
    return np.load(os.path.join(GOLDEN, "lethal.npz"))


@pytest.fixture(scope="session")
def golden_batch():
    return np.load(os.path.join(GOLDEN, "batch.npz"))
