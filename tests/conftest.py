import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_pkg(sub: str = ""):
    """The product package directory is 'gd-slam_b200' (hyphen): import it through importlib."""
    name = "gd-slam_b200" + (("." + sub) if sub else "")
    return importlib.import_module(name)


@pytest.fixture(scope="session")
def synth():
    return load_pkg("synth")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle

    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def golden():
    def _load(name):
        return np.load(os.path.join(GOLDEN, name))

    return _load


def flow_tol_violations(a, ref):
    """BASELINE.md tolerance for flow / Mahalanobis values: |d| <= 1e-4 * max(1, |ref|)."""
    d = np.abs(a.astype(np.float64) - ref.astype(np.float64))
    return int((d > 1e-4 * np.maximum(1.0, np.abs(ref))).sum()), float(d.max())
