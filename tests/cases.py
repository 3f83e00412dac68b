"""Seeded layer cases shared by the golden generator (tests/golden/make_golden.py, run against the
real reference in the build container) and by the CPU / GPU parity tests."""
import numpy as np

# name, kind, shapes and hyper-parameters.  x: input shape; w: weight shape.
CASES = [
    dict(name="baseline_cfg1_c64_32x32", kind="conv2d", x=(1, 64, 32, 32), w=(64, 64, 3, 3), pad=(1, 1), center=True),
    dict(name="s2_relu_fed", kind="conv2d", x=(2, 64, 14, 14), w=(128, 64, 3, 3), stride=(2, 2), pad=(1, 1), relu=True, center=True),
    dict(name="k1_plain", kind="conv2d", x=(2, 64, 7, 7), w=(128, 64, 1, 1)),
    dict(name="bias_post_relu", kind="conv2d", x=(2, 128, 9, 11), w=(64, 128, 3, 3), pad=(1, 1), bias=True, post=True, relu=True, center=True),
    dict(name="ragged_c3", kind="conv2d", x=(2, 3, 8, 8), w=(16, 3, 1, 1), bias=True),
    dict(name="ragged_c70_s21", kind="conv2d", x=(2, 70, 6, 7), w=(40, 70, 3, 3), stride=(2, 1), pad=(1, 1), bias=True, post=True, center=True, relu=True),
    dict(name="dilated", kind="conv2d", x=(1, 64, 12, 12), w=(32, 64, 3, 3), pad=(2, 2), dil=(2, 2), center=True),
    dict(name="k5", kind="conv2d", x=(1, 32, 10, 10), w=(48, 32, 5, 5), pad=(2, 2), relu=True),
    dict(name="k1x7", kind="conv2d", x=(1, 64, 5, 20), w=(64, 64, 1, 7), pad=(0, 3), center=True),
    dict(name="no_alpha", kind="conv2d", x=(2, 64, 8, 8), w=(64, 64, 3, 3), pad=(1, 1), alpha=False),
    dict(name="deep_7x7", kind="conv2d", x=(3, 256, 7, 7), w=(256, 256, 3, 3), pad=(1, 1), center=True, relu=True),
    dict(name="chunks8", kind="conv2d", x=(2, 512, 7, 7), w=(64, 512, 3, 3), pad=(1, 1), center=True),
    dict(name="s2_odd", kind="conv2d", x=(1, 128, 15, 15), w=(128, 128, 3, 3), stride=(2, 2), pad=(1, 1), center=True, relu=True),
    dict(name="k1_s2", kind="conv2d", x=(2, 64, 8, 8), w=(128, 64, 1, 1), stride=(2, 2)),
    dict(name="wide_300", kind="conv2d", x=(1, 64, 3, 300), w=(32, 64, 3, 3), pad=(1, 1), relu=True),
    dict(name="valid_pad0", kind="conv2d", x=(2, 64, 9, 9), w=(96, 64, 3, 3), center=True, post=True),
    dict(name="w28_c128", kind="conv2d", x=(1, 128, 28, 28), w=(128, 128, 3, 3), pad=(1, 1), center=True, relu=True),
    dict(name="groups2", kind="conv2d", x=(2, 64, 9, 9), w=(64, 32, 3, 3), pad=(1, 1), groups=2, center=True, bias=True, post=True),
    dict(name="groups12_dil2", kind="conv2d", x=(1, 96, 10, 10), w=(96, 8, 3, 3), pad=(2, 2), dil=(2, 2), groups=12, relu=True),
    dict(name="lin_small", kind="linear", x=(5, 100), w=(10, 100), bias=True, post=True),
    dict(name="lin_fc", kind="linear", x=(64, 512), w=(1000, 512), bias=True, center=True),
    dict(name="lin_row1", kind="linear", x=(1, 64), w=(64, 64)),
    dict(name="lin_lead_dims", kind="linear", x=(2, 3, 128), w=(32, 128), bias=True, relu=True),
    dict(name="c1d_k3", kind="conv1d", x=(2, 64, 50), w=(64, 64, 3), pad=(1,), center=True),
    dict(name="c1d_k5_s2", kind="conv1d", x=(2, 16, 33), w=(8, 16, 5), stride=(2,), pad=(2,), bias=True, post=True),
]


def by_name(name):
    return next(c for c in CASES if c["name"] == name)


def make_inputs(case, seed_offset=0):
    """Deterministic inputs: standard-normal activations (optionally ReLU-ed => ~50 % exact zeros,
    plus a few -0.0), kaiming-like weights, random bias / positive post scale."""
    idx = [c["name"] for c in CASES].index(case["name"])
    rng = np.random.default_rng(1000 + idx + 7919 * seed_offset)
    x = rng.standard_normal(case["x"]).astype(np.float32)
    if case.get("relu"):
        x = np.maximum(x, 0.0).astype(np.float32)
        flat = x.reshape(-1)
        flat[:: 17][flat[:: 17] == 0] = -0.0
    w = (rng.standard_normal(case["w"]) * 0.05).astype(np.float32)
    c_out = case["w"][0]
    bias = (rng.standard_normal(c_out) * 0.5).astype(np.float32) if case.get("bias") else None
    post = (0.5 + rng.random(c_out)).astype(np.float32) if case.get("post") else None
    return x, w, bias, post


def hyper(case):
    nd = {"conv2d": 2, "conv1d": 1, "linear": 0}[case["kind"]]
    one = (1,) * nd
    return dict(stride=tuple(case.get("stride", one)), pad=tuple(case.get("pad", (0,) * nd)),
                dil=tuple(case.get("dil", one)), center=bool(case.get("center", False)),
                alpha=bool(case.get("alpha", True)), groups=int(case.get("groups", 1)))
