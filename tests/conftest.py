import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_layers():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "layers.npz"))


@pytest.fixture(scope="session")
def golden_units():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "ref_unit_vectors.npz"))


@pytest.fixture(scope="session")
def golden_models():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "models.npz"))


def rel_err(a, b):
    import numpy as np
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
