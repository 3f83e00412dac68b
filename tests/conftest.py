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


def pytest_collection_modifyitems(config, items):
    """A plain ``pytest`` on a CUDA-less host skips the GPU tests instead of failing at the first one.  An explicit
    ``-m gpu`` run (the B200 box) never skips: there a missing device or library must be a loud failure."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    import torch
    from bnn_b200 import native
    if torch.cuda.is_available() and native.available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and the built libbnn_b200.so")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
