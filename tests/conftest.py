import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_small():
    return np.load(os.path.join(GOLDEN, "reference_small.npz"))


@pytest.fixture(scope="session")
def golden_configs():
    return np.load(os.path.join(GOLDEN, "reference_configs.npz"))


@pytest.fixture(scope="session")
def golden_classes():
    return np.load(os.path.join(GOLDEN, "reference_classes.npz"))


@pytest.fixture(scope="session")
def c2_data(golden_configs):
    """make_regression(10000, 100, random_state=0): BASELINE.json configs[0], configs[1]."""
    import hashlib
    from sklearn.datasets import make_regression
    X, _ = make_regression(n_samples=10000, n_features=100, random_state=0)
    if hashlib.sha256(np.ascontiguousarray(X).tobytes()).hexdigest() != str(golden_configs["X_sha"]):
        pytest.skip("sklearn here generates different make_regression data than the golden run")
    # y = X @ coef is computed by BLAS and differs in the last bits between CPUs (which flips
    # exact-tie decisions in tiny nodes), so the golden run's y is stored in the fixture
    return X, golden_configs["y"]
