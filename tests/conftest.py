import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = "n-body_pointcloudevolution_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def syn():
    return importlib.import_module(PKG + ".synthetic")


@pytest.fixture(scope="session")
def nb():
    """The product package (imports the CUDA C-ABI library)."""
    return importlib.import_module(PKG)
