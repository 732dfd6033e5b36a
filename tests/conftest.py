import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def rel_l2(a, b):
    import numpy as np
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_propagation.npz"))


def check(name, err, tol):
    """Assert err < tol and print the measured figure (pytest -s) so that tolerances can be audited."""
    print(f"PARITY {name}: {err:.3e} (tol {tol:.0e})")
    assert err < tol, (name, err, tol)


def rel_scalar(a, b):
    return abs(float(a) - float(b)) / max(abs(float(b)), 1e-300)
