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


@pytest.fixture(scope="session")
def golden_cfg34():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "models_cfg34.npz"))


def build_config34(which):
    """BASELINE configs[2] / configs[3] prepared with bnn_b200, seeded exactly like tests/golden/make_golden.py builds them
    on the real reference; returns (prepared eval model on CPU, input [2,3,res,res])."""
    import torch
    import bnn_b200 as bnn
    from bnn_b200 import workloads
    from bnn_b200.ops import BasicInputBinarizer, BasicScaleBinarizer, XNORWeightBinarizer
    torch.manual_seed(0)
    if which == "resnet50":
        m, post, res = workloads.resnet50(), BasicScaleBinarizer, 96
    else:
        m, post, res = workloads.HBlockNet(), bnn.Identity, 64
    cfg = bnn.BConfig(BasicInputBinarizer, post, XNORWeightBinarizer.with_args(compute_alpha=True, center_weights=True))
    m = bnn.prepare_binary_model(m, cfg, ignore_layers_name=["_first_", "_last_"])
    workloads.randomize_batchnorm(m, seed=1)
    x = torch.randn(2, 3, res, res, generator=torch.Generator().manual_seed(0))
    return m.eval(), x


def rel_err(a, b):
    import numpy as np
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
