"""Generate the committed golden fixtures from the REAL reference (read-only at /root/reference).

Run in the build container only:   PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
Outputs (small .npz files next to this script):
  ref_unit_vectors.npz  the reference's own known-answer vectors (test/test_layers.py:22-66,
                        test/test_binarize.py:118-120) together with what the reference returns
  layers.npz            reference CPU-fp32 outputs for every case of tests/cases.py
  models.npz            logits of reference ResNet-18 models (64x64 inputs, randomised BN) + a
                        parameter checksum so the seeded re-construction can be verified
  models_cfg34.npz      the same for BASELINE configs[2] (ResNet-50 XNOR-Net++, 96x96) and configs[3] (the
                        Hierarchical-Block harness around the reference's HBlock, 64x64)
Nothing here is imported at test time; tests only read the .npz files.
"""
import importlib
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REF = "/root/reference/bnn"
spec = importlib.util.spec_from_file_location("bnn_ref", os.path.join(REF, "__init__.py"),
                                              submodule_search_locations=[REF])
bnn_ref = importlib.util.module_from_spec(spec)
sys.modules["bnn_ref"] = bnn_ref
spec.loader.exec_module(bnn_ref)
ref_ops = importlib.import_module("bnn_ref.ops")
ref_resnet = importlib.import_module("bnn_ref.models.resnet")
ref_blocks = importlib.import_module("bnn_ref.models.layers")

import cases  # noqa: E402
import bnn_b200  # noqa: E402  (only for workloads.randomize_batchnorm / checksum helpers)
from bnn_b200 import workloads  # noqa: E402

torch.set_grad_enabled(False)


def ref_layer(case):
    x, w, bias, post = cases.make_inputs(case)
    hp = cases.hyper(case)
    wb = ref_ops.XNORWeightBinarizer.with_args(compute_alpha=hp["alpha"], center_weights=hp["center"])
    cfg = bnn_ref.BConfig(activation_pre_process=ref_ops.BasicInputBinarizer,
                          activation_post_process=ref_ops.BasicScaleBinarizer if post is not None else bnn_ref.Identity,
                          weight_pre_process=wb)
    if case["kind"] == "conv2d":
        m = nn.Conv2d(w.shape[1] * hp["groups"], w.shape[0], w.shape[2:], stride=hp["stride"], padding=hp["pad"],
                      dilation=hp["dil"], groups=hp["groups"], bias=bias is not None)
    elif case["kind"] == "conv1d":
        m = nn.Conv1d(w.shape[1], w.shape[0], w.shape[2], stride=hp["stride"], padding=hp["pad"],
                      dilation=hp["dil"], bias=bias is not None)
    else:
        m = nn.Linear(w.shape[1], w.shape[0], bias=bias is not None)
    m.weight.data.copy_(torch.from_numpy(w))
    if bias is not None:
        m.bias.data.copy_(torch.from_numpy(bias))
    m = bnn_ref.prepare_binary_model(m, cfg)
    if post is not None:
        m.activation_post_process.alpha.data.copy_(torch.from_numpy(post).reshape(m.activation_post_process.alpha.shape))
    return m(torch.from_numpy(x)).numpy()


def unit_vectors():
    data = torch.tensor([-0.05263, -0.05068, -0.03849, 0.03104, 0.0772, 0.03038, -0.06640, 0.05894,
                         0.13059, 0.03433, -0.25811, 0.13785]).view(1, 3, 2, 2)
    weights = torch.tensor([-0.0252, 0.0084, -0.0676, 0.0891, -0.0010, 0.0518, 0.0380, 0.2866, -0.0050])
    cfg = bnn_ref.BConfig(activation_pre_process=ref_ops.BasicInputBinarizer,
                          activation_post_process=ref_ops.BasicScaleBinarizer,
                          weight_pre_process=ref_ops.XNORWeightBinarizer)
    out = {"data": data.numpy(), "weights": weights.numpy()}
    lin = nn.Linear(3, 3, bias=False); lin.weight.data.copy_(weights.view(3, 3))
    out["linear_ref"] = bnn_ref.prepare_binary_model(lin, cfg)(data[:, :, 0, 0].view(1, 3)).numpy()
    out["linear_expected"] = np.array([[0.0337, -0.0473, -0.1099]], np.float32)
    c1 = nn.Conv1d(3, 3, 1, bias=False); c1.weight.data.copy_(weights.view(3, 3, 1))
    out["conv1d_ref"] = bnn_ref.prepare_binary_model(c1, cfg)(data[:, :, :, 0].reshape(1, 3, 2)).numpy()
    out["conv1d_expected"] = np.array([[[0.0337, 0.0337], [-0.0473, -0.0473], [-0.1099, -0.1099]]], np.float32)
    c2 = nn.Conv2d(3, 3, 1, bias=False); c2.weight.data.copy_(weights.view(3, 3, 1, 1))
    out["conv2d_ref"] = bnn_ref.prepare_binary_model(c2, cfg)(data).numpy()
    out["conv2d_expected"] = np.array([[[[0.0337, 0.0337], [0.0337, -0.0337]], [[-0.0473, -0.0473], [-0.0473, 0.0473]],
                                        [[-0.1099, -0.1099], [-0.1099, 0.1099]]]], np.float32)
    s_in = torch.tensor([0.3, 0.1, -2, -0.001, 0.01, 0.0, -0.0, 1e-30, -1e-30, float("nan")])
    out["sign_in"] = s_in.numpy()
    out["sign_ref"] = ref_ops.BasicInputBinarizer()(s_in.clone()).numpy()
    return out


def param_checksum(model):
    return np.array([float(p.detach().double().abs().sum()) for p in model.state_dict().values()
                     if p.dtype.is_floating_point], np.float64)


def model_logits(variant):
    torch.manual_seed(0)
    if variant == "basic_relu":
        m = ref_resnet.resnet18()
    else:
        m = ref_resnet.resnet18(block_type=ref_blocks.PreBasicBlock, activation=nn.PReLU)
    cfg = bnn_ref.BConfig(activation_pre_process=ref_ops.BasicInputBinarizer,
                          activation_post_process=bnn_ref.Identity,
                          weight_pre_process=ref_ops.XNORWeightBinarizer.with_args(compute_alpha=True, center_weights=True))
    m = bnn_ref.prepare_binary_model(m, cfg, ignore_layers_name=["_first_", "_last_"])
    workloads.randomize_batchnorm(m, seed=1)
    m.eval()
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    return m(x).numpy(), param_checksum(m), m.double()(x.double()).numpy()


def config34_logits(which):
    """BASELINE configs[2] / configs[3] on the REAL reference: ResNet-50 XNOR-Net++ (fc patched to 2048 inputs, learned
    per-channel post scale randomised) and the Hierarchical-Block harness built around the reference's own HBlock."""
    torch.manual_seed(0)
    if which == "resnet50":
        m = ref_resnet.resnet50()
        m.fc = nn.Linear(2048, 1000)                       # upstream wires fc to 512 features (resnet.py:101,143)
        post, res = ref_ops.BasicScaleBinarizer, 96
    else:
        m = workloads.HBlockNet(hblock=lambda i, p, d: ref_blocks.HBlock(i, p, downsample=d, norm_layer=nn.BatchNorm2d))
        post, res = bnn_ref.Identity, 64
    cfg = bnn_ref.BConfig(activation_pre_process=ref_ops.BasicInputBinarizer, activation_post_process=post,
                          weight_pre_process=ref_ops.XNORWeightBinarizer.with_args(compute_alpha=True, center_weights=True))
    m = bnn_ref.prepare_binary_model(m, cfg, ignore_layers_name=["_first_", "_last_"])
    workloads.randomize_batchnorm(m, seed=1)
    m.eval()
    x = torch.randn(2, 3, res, res, generator=torch.Generator().manual_seed(0))
    return m(x).numpy(), param_checksum(m), m.double()(x.double()).numpy()


def main():
    cfg34 = {}
    for which in ("resnet50", "hblock"):
        logits, chk, logits64 = config34_logits(which)
        cfg34[which + "_logits"], cfg34[which + "_checksum"] = logits, chk
        cfg34[which + "_logits_fp64"] = logits64.astype(np.float64)
        print(which, logits.shape, "fp32 vs fp64 of the reference itself:", np.abs(logits - logits64).max() / np.abs(logits64).max())
    np.savez_compressed(os.path.join(HERE, "models_cfg34.npz"), **cfg34)
    np.savez_compressed(os.path.join(HERE, "ref_unit_vectors.npz"), **unit_vectors())
    layers = {}
    for case in cases.CASES:
        layers[case["name"]] = ref_layer(case)
        print(case["name"], layers[case["name"]].shape)
    np.savez_compressed(os.path.join(HERE, "layers.npz"), **layers)
    models = {}
    for variant in ("basic_relu", "pre_prelu"):
        logits, chk, logits64 = model_logits(variant)
        models[variant + "_logits"] = logits
        models[variant + "_checksum"] = chk
        models[variant + "_logits_fp64"] = logits64.astype(np.float64)
        print(variant, np.abs(logits - logits64).max() / np.abs(logits64).max())
    np.savez_compressed(os.path.join(HERE, "models.npz"), **models)


if __name__ == "__main__":
    main()
