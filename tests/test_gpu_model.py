"""Whole-model parity and full-size properties on the GPU."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import rel_err
from oracle import floatsim as fs

import bnn_b200 as bnn
from bnn_b200 import functional as BF
from bnn_b200 import native, workloads
from bnn_b200.ops import BasicInputBinarizer, BasicScaleBinarizer, XNORWeightBinarizer

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def xnor_cfg(post=bnn.Identity):
    return bnn.BConfig(BasicInputBinarizer, post, XNORWeightBinarizer.with_args(compute_alpha=True, center_weights=True))


def build(variant, cfg=None):
    torch.manual_seed(0)
    m = workloads.resnet18() if variant == "basic_relu" else workloads.resnet18(workloads.PreBasicBlock, nn.PReLU)
    m = bnn.prepare_binary_model(m, cfg or xnor_cfg(), ignore_layers_name=["_first_", "_last_"])
    workloads.randomize_batchnorm(m, seed=1)
    return m.eval()


@pytest.fixture(autouse=True)
def _fp32_glue():
    # the non-binarized glue (stem conv, fc) must not run in TF32, or parity is lost there
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("variant", ["basic_relu", "pre_prelu"])
def test_resnet18_logits_match_the_reference(variant, golden_models):
    m = build(variant).to(DEV)
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    before = native.launch_count()
    with torch.no_grad():
        y = m(x.to(DEV)).cpu().numpy()
    assert native.launch_count() - before >= 2 * 19
    ref = golden_models[variant + "_logits"]
    # north_star: within 1e-3 relative of the reference's float-sim forward
    assert rel_err(y, ref) <= 1e-3, rel_err(y, ref)
    assert (np.argmax(y, 1) == np.argmax(ref, 1)).all()


def test_resnet18_full_resolution_vs_oracle_twin():
    """224x224, batch 4: CUDA engine vs the oracle's CPU float simulation of the same prepared model,
    with a learned XNOR-Net++ post scale so the fused epilogue scale is exercised end to end."""
    m = build("basic_relu", xnor_cfg(BasicScaleBinarizer))
    twin = fs.mirror_model(m)
    x = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        want = twin(x).numpy()
        got = m.to(DEV)(x.to(DEV)).cpu().numpy()
    assert rel_err(got, want) <= 1e-3, rel_err(got, want)


def test_full_size_layer_properties():
    """BASELINE-size layer (bs 256, 64ch, 56x56) through size-independent identities:
    batch-split invariance, antisymmetry in x, zero input, integer parity of the dot."""
    torch.manual_seed(0)
    n = 256
    x = torch.relu(torch.randn(n, 64, 56, 56, device=DEV))          # ~50 % exact zeros, like a ReLU-fed layer
    w = torch.randn(64, 64, 3, 3, device=DEV) * 0.05
    wts = BF.pack_weights(w, True, True)
    act = BF.pack_activations(x)
    dot = BF.bconv2d(act, wts, None, None, (1, 1), (1, 1), (1, 1), use_alpha=False)
    # (a) any sub-batch gives bit-identical rows
    sub = BF.bconv2d(BF.pack_activations(x[100:132]), wts, None, None, (1, 1), (1, 1), (1, 1), use_alpha=False)
    assert torch.equal(dot[100:132], sub)
    # (b) dot(-x) == -dot(x) exactly (mask plane unchanged, sign plane flipped)
    neg = BF.bconv2d(BF.pack_activations(-x[:16]), wts, None, None, (1, 1), (1, 1), (1, 1), use_alpha=False)
    assert torch.equal(neg, -dot[:16])
    # (c) dot and the number of non-zero inputs under the window have the same parity
    ones = torch.ones(1, 1, 3, 3, device=DEV)
    nz = torch.nn.functional.conv2d((x[:16] != 0).float().sum(1, keepdim=True), ones, padding=1)
    assert torch.equal(torch.remainder(dot[:16], 2), torch.remainder(nz, 2).expand(-1, 64, -1, -1))
    assert dot.abs().max().item() <= 576
    # (d) all-zero input -> bias * post exactly
    bias = torch.randn(64, device=DEV)
    post = torch.rand(64, device=DEV) + 0.5
    z = BF.bconv2d(BF.pack_activations(torch.zeros(2, 64, 56, 56, device=DEV)), wts, bias, post, (1, 1), (1, 1), (1, 1))
    assert torch.equal(z, (bias * post).view(1, -1, 1, 1).expand_as(z))
    # (e) epilogue is exactly (alpha*dot + bias)*post in fp32
    y = BF.bconv2d(BF.pack_activations(x[:8]), wts, bias, post, (1, 1), (1, 1), (1, 1))
    want = (wts.alpha.view(1, -1, 1, 1) * dot[:8] + bias.view(1, -1, 1, 1)) * post.view(1, -1, 1, 1)
    assert torch.equal(y, want)


def test_full_batch_model_rows_are_independent():
    m = build("basic_relu").to(DEV)
    x = torch.randn(64, 3, 224, 224, device=DEV)
    with torch.no_grad():
        full = m(x)
        part = m(x[16:24])
    assert torch.isfinite(full).all()
    assert rel_err(full[16:24].cpu().numpy(), part.cpu().numpy()) <= 1e-5


def test_forward_is_cuda_graph_capturable():
    m = build("basic_relu").to(DEV)
    x = torch.randn(8, 3, 96, 96, device=DEV)
    with torch.no_grad():
        eager = m(x)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            m(x)
            torch.cuda.current_stream().synchronize()
            with torch.cuda.graph(g, stream=s):
                out = m(x)
        g.replay()
        torch.cuda.synchronize()
    assert torch.equal(out, eager)


def test_baseline_config3_resnet50_xnor_plus_plus():
    """BASELINE configs[2]: ResNet-50 (fc patched to 2048), XNOR-Net++ recipe = learned BasicScaleBinarizer
    (examples/recepies/xnor-net-plus.yaml:13-25), 52 binarized convs, vs the oracle's float-simulated twin."""
    torch.manual_seed(0)
    m = workloads.resnet50()
    m = bnn.prepare_binary_model(m, xnor_cfg(BasicScaleBinarizer), ignore_layers_name=["_first_", "_last_"])
    assert sum(isinstance(x, bnn.layers.Conv2d) for x in m.modules()) == 52
    workloads.randomize_batchnorm(m, seed=1)
    m.eval()
    twin = fs.mirror_model(m)
    x = torch.randn(2, 3, 96, 96, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        want = twin(x).numpy()
        got = m.to(DEV)(x.to(DEV)).cpu().numpy()
    assert rel_err(got, want) <= 1e-3, rel_err(got, want)


def test_baseline_config4_hierarchical_block_net():
    """BASELINE configs[3]: Hierarchical-Block harness, ReLU-fed => {0,+1} activations, 16
    binarized 3x3/1x1 convs incl. 128->64 and 64->64 (C=2 / C=1 channel tiles), vs the float-simulated twin."""
    torch.manual_seed(0)
    m = workloads.HBlockNet(depth=2)
    m = bnn.prepare_binary_model(m, xnor_cfg(), ignore_layers_name=["_first_", "_last_"])
    workloads.randomize_batchnorm(m, seed=1)
    m.eval()
    assert sum(isinstance(x, bnn.layers.Conv2d) for x in m.modules()) == 10
    twin = fs.mirror_model(m)
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        want = twin(x).numpy()
        got = m.to(DEV)(x.to(DEV)).cpu().numpy()
    assert rel_err(got, want) <= 1e-3, rel_err(got, want)


@pytest.mark.parametrize("which", ["resnet50", "hblock"])
def test_config34_logits_match_the_real_reference(which, golden_cfg34):
    """BASELINE configs[2] / [3] against logits of the REAL reference (tests/golden/models_cfg34.npz, generated by
    tests/golden/make_golden.py from /root/reference): per-layer path and fused engine, tolerance 1e-3 (north_star)."""
    from conftest import build_config34
    from bnn_b200 import fuse
    m, x = build_config34(which)
    ref = golden_cfg34[which + "_logits"]
    m = m.to(DEV)
    with torch.no_grad():
        per_layer = m(x.to(DEV)).cpu().numpy()
        engine = fuse.optimize(m)
        assert engine is not m
        fused = engine(x.to(DEV)).cpu().numpy()
    print(which, "per-layer vs reference", rel_err(per_layer, ref), "fused vs reference", rel_err(fused, ref))
    assert rel_err(per_layer, ref) <= 1e-3 and rel_err(fused, ref) <= 1e-3
    assert (np.argmax(per_layer, 1) == np.argmax(ref, 1)).all() and (np.argmax(fused, 1) == np.argmax(ref, 1)).all()


def test_reference_resnet_with_dabnn_stem_through_the_reference_entry_point():
    """the reference's own ResNet with ``stem_type='dabnn'`` (bnn/models/resnet.py:10-47), built from the
    byte-compiled reference (oracle/_ref), converted by the REFERENCE's prepare_binary_model with our module mapping.
    The fused engine keeps the DaBNN stem as torch ops and fuses the eight residual blocks; its logits match the same
    model with the reference's own float-simulated layers on the CPU."""
    from oracle import build as oracle_build
    ref = oracle_build.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    import importlib
    from bnn_b200 import fuse
    rops = importlib.import_module("bnn_ref.ops")
    rres = importlib.import_module("bnn_ref.models.resnet")
    rcfg = ref.BConfig(activation_pre_process=rops.BasicInputBinarizer, activation_post_process=ref.Identity,
                       weight_pre_process=rops.XNORWeightBinarizer.with_args(compute_alpha=True, center_weights=True))
    torch.manual_seed(7)
    plain = rres.resnet18(stem_type="dabnn")
    workloads.randomize_batchnorm(plain, seed=2)
    state = {k: v.clone() for k, v in plain.state_dict().items()}
    ours = ref.prepare_binary_model(plain, rcfg, modules_mapping=bnn.mapping_for_reference(ref),
                                    ignore_layers_name=["_first_", "_last_"]).eval().to(DEV)
    theirs = rres.resnet18(stem_type="dabnn")
    theirs.load_state_dict(state)
    theirs = ref.prepare_binary_model(theirs, rcfg, ignore_layers_name=["_first_", "_last_"]).eval()
    engine = fuse.optimize(ours)
    assert isinstance(engine, fuse.FusedResNet) and engine.fused_blocks == 8 and not engine.stem.ok
    x = torch.randn(2, 3, 96, 96, generator=torch.Generator().manual_seed(11))
    with torch.no_grad():
        want = theirs(x).numpy()
        got = engine(x.to(DEV)).cpu().numpy()
        again = engine(x.to(DEV)).cpu().numpy()            # second forward: shortcuts on the second stream
        per_layer = ours(x.to(DEV)).cpu().numpy()
    assert engine.stem_kernel_used == "torch"
    assert np.array_equal(got, again)
    assert rel_err(per_layer, want) <= 1e-3
    assert rel_err(got, want) <= 1e-3
