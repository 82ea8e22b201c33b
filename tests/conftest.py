import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import htsp_b200  # noqa: E402,F401  (registers the package from h-tsp_b200/)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "path_solver_golden.npz"))


@pytest.fixture(scope="session")
def built_lib():
    """Builds (or reuses) libhtsp_b200.so; nvcc cross-compiles sm_100a without a GPU."""
    from htsp_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def sd6():
    from htsp_b200.weights import synthetic_state_dict
    return synthetic_state_dict(1234, 6)
