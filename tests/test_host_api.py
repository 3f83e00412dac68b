"""Host-side mirror of the reference interface: conversion rules, module protocol, state dict,
error conventions (reference test/test_binarize.py, test/test_layers.py re-stated for bnn_b200).
CPU only; forwards use the explicit float-simulation opt-in."""
import copy
import logging
import os

import pytest
import torch
import torch.nn as nn

import bnn_b200 as bnn
from bnn_b200 import runtime, workloads
from bnn_b200.layers import Conv1d, Conv2d, Linear
from bnn_b200.native import NativeError
from bnn_b200.ops import (AdvancedInputBinarizer, BasicInputBinarizer, BasicScaleBinarizer, SignActivation,
                          StochasticInputBinarizer, XNORScaleBinarizer, XNORWeightBinarizer)

CFG = bnn.BConfig(activation_pre_process=BasicInputBinarizer, activation_post_process=BasicScaleBinarizer,
                  weight_pre_process=XNORWeightBinarizer)


def small_net():
    return nn.Sequential(nn.Conv2d(3, 16, 1, 1), nn.BatchNorm2d(16), nn.ReLU(inplace=True), nn.Conv2d(16, 16, 1, 1),
                         nn.BatchNorm2d(16), nn.ReLU(inplace=True), nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten(),
                         nn.Linear(16, 3))


def test_single_leaf_modules_are_returned_converted():
    assert type(bnn.prepare_binary_model(nn.Linear(10, 3), CFG)) is Linear
    assert type(bnn.prepare_binary_model(nn.Conv2d(3, 16, 1), CFG)) is Conv2d
    assert type(bnn.prepare_binary_model(nn.Conv1d(3, 16, 1), CFG)) is Conv1d


def test_parameters_are_shared_not_copied():
    src = nn.Conv2d(4, 8, 3, stride=2, padding=1, dilation=1, bias=True)
    new = Conv2d.from_module(src, CFG)
    assert new.weight is src.weight and new.bias is src.bias
    assert (new.stride, new.padding, new.kernel_size) == (src.stride, src.padding, src.kernel_size)
    assert isinstance(new.activation_pre_process, BasicInputBinarizer)
    assert isinstance(new.activation_post_process, BasicScaleBinarizer)
    assert isinstance(new.weight_pre_process, XNORWeightBinarizer)
    assert new.bconfig is CFG
    assert list(new.state_dict().keys()) == ["weight", "bias", "activation_post_process.alpha"]
    assert new.activation_post_process.alpha.shape == (1, 8, 1, 1)
    assert Linear.from_module(nn.Linear(5, 7), CFG).activation_post_process.alpha.shape == (1, 7)


def test_from_module_contract_errors():
    with pytest.raises(AssertionError):
        Conv2d.from_module(nn.Linear(3, 3), CFG)
    with pytest.raises(AssertionError):
        Conv2d.from_module(nn.Conv2d(3, 3, 1))          # float module has no bconfig
    with pytest.raises(AssertionError):
        Conv2d(3, 3, 1)                                 # bconfig required
    again = Conv2d.from_module(Conv2d.from_module(nn.Conv2d(3, 3, 1), CFG))   # re-convert takes mod.bconfig
    assert again.bconfig is CFG


def test_bconfig_rejects_instances_and_with_args_chains():
    with pytest.raises(ValueError):
        bnn.BConfig(activation_pre_process=BasicInputBinarizer())
    f = XNORWeightBinarizer.with_args(compute_alpha=False).with_args(center_weights=True)
    a, b = f(), f()
    assert a is not b and (a.compute_alpha, a.center_weights) == (False, True)
    assert XNORWeightBinarizer.with_args(compute_alpha=False).with_args(compute_alpha=True)().compute_alpha is True
    bnn.BConfig(weight_pre_process=f)                  # factories are fine
    with pytest.raises(ValueError):
        XNORWeightBinarizer()._compute_alpha(torch.zeros(2, 2, 2, 2, 2))


def test_whole_net_conversion_and_per_layer_override():
    fp32 = bnn.BConfig(nn.Identity, nn.Identity, nn.Identity)
    model = bnn.prepare_binary_model(small_net(), CFG, custom_config_layers_name={"8": fp32})
    convs = [m for m in model.modules() if isinstance(m, Conv2d)]
    lins = [m for m in model.modules() if isinstance(m, Linear)]
    assert len(convs) == 2 and len(lins) == 1
    assert isinstance(lins[0].activation_pre_process, nn.Identity)
    assert isinstance(lins[0].weight_pre_process, nn.Identity)


def test_ignore_rules_special_words_regex_literal():
    names = lambda m: [n for n, x in m.named_modules() if isinstance(x, (Conv2d, Linear))]
    assert names(bnn.prepare_binary_model(small_net(), CFG, ignore_layers_name=["_first_", "_last_"])) == ["3"]
    # upstream binds the words crosswise (binarize.py:47-50): '_first_' alone skips the LAST layer
    assert names(bnn.prepare_binary_model(small_net(), CFG, ignore_layers_name=["_first_"])) == ["0", "3"]
    assert names(bnn.prepare_binary_model(small_net(), CFG, ignore_layers_name=["_last_"])) == ["3", "8"]
    assert names(bnn.prepare_binary_model(small_net(), CFG, ignore_layers_name=["$^[03]$$"])) == ["8"]
    assert names(bnn.prepare_binary_model(small_net(), CFG, ignore_layers_name=["3"])) == ["0", "8"]
    torch.manual_seed(0)
    r18 = bnn.prepare_binary_model(workloads.resnet18(), CFG, ignore_layers_name=["_first_", "_last_"])
    assert type(r18.conv1) is nn.Conv2d and type(r18.fc) is nn.Linear
    assert sum(isinstance(m, Conv2d) for m in r18.modules()) == 19
    assert isinstance(r18.layer2[0].downsample[1], Conv2d)


def test_unknown_custom_name_warns_only_for_existing_modules(caplog):
    with caplog.at_level(logging.WARNING):
        bnn.prepare_binary_model(small_net(), CFG, custom_config_layers_name={"1": CFG})   # '1' is a BatchNorm
    assert "Module named 1" in caplog.text


def test_multi_device_module_is_rejected():
    m = nn.Conv2d(3, 3, 1)
    m.register_buffer("meta_buf", torch.empty(1, device="meta"))
    with pytest.raises(AssertionError):
        bnn.get_modules_to_binarize(nn.Sequential(m), CFG)


def test_default_forward_never_falls_back_silently():
    layer = bnn.prepare_binary_model(nn.Conv2d(64, 64, 3, padding=1), CFG).eval()
    x = torch.randn(1, 64, 8, 8)
    with torch.no_grad(), pytest.raises(NativeError, match="no CPU implementation"):
        layer(x)
    with runtime.floatsim_enabled(), torch.no_grad():
        assert layer(x).shape == (1, 64, 8, 8)
    assert runtime.floatsim() is False


def test_training_mode_routes_to_the_float_simulation_with_one_warning():
    """reference bnn/layers/conv.py:90-97 runs under autograd; a prepared bnn_b200 module in train() mode does the same
    through the torch float simulation (STE backward, bnn/ops.py:68-73) and says so once."""
    import warnings as w
    from bnn_b200.layers import binary_modules
    binary_modules._warned_training = False
    layer = bnn.prepare_binary_model(nn.Conv2d(64, 64, 3, padding=1), CFG)          # default state: training
    x = torch.randn(2, 64, 8, 8)
    with pytest.warns(RuntimeWarning, match="fp32 simulation"):
        y = layer(x)
    assert y.grad_fn is not None
    y.sum().backward()
    assert layer.weight.grad is not None and torch.isfinite(layer.weight.grad).all()
    with w.catch_warnings():
        w.simplefilter("error")                        # second call: no warning any more
        layer(x)
    # same numbers as the explicit opt-in / the reference's forward
    with runtime.floatsim_enabled(), torch.no_grad():
        assert torch.equal(layer.eval()(x), y.detach())
    # eval-mode CPU tensors still raise: no silent fallback outside training
    with torch.no_grad(), pytest.raises(NativeError, match="no CPU implementation"):
        layer(x)


def test_runtime_switches_are_process_wide_with_thread_local_override():
    import threading
    seen = {}
    runtime.floatsim(True)
    try:
        t = threading.Thread(target=lambda: seen.setdefault("worker", runtime.floatsim()))
        t.start(); t.join()
        assert seen["worker"] is True                  # nn.DataParallel replica threads see the main thread's opt-in
        with runtime.floatsim_enabled(False):
            assert runtime.floatsim() is False
            t = threading.Thread(target=lambda: seen.setdefault("worker2", runtime.floatsim()))
            t.start(); t.join()
            assert seen["worker2"] is True             # the override is per thread
        assert runtime.floatsim() is True
    finally:
        runtime.floatsim(False)
    assert runtime.floatsim() is False
    assert runtime.shortcut_max_cin() == 1024          # engine tuning knob: one-kernel shortcuts up to 1024 input channels
    try:
        assert runtime.shortcut_max_cin(128) == 128
    finally:
        runtime.shortcut_max_cin(1024)


def test_floatsim_matches_reference_known_answers(golden_units):
    g = golden_units
    w = torch.from_numpy(g["weights"])
    with runtime.floatsim_enabled(), torch.no_grad():
        lin = nn.Linear(3, 3, bias=False); lin.weight.data.copy_(w.view(3, 3))
        out = bnn.prepare_binary_model(lin, CFG)(torch.from_numpy(g["data"])[:, :, 0, 0].reshape(1, 3))
        assert torch.allclose(out, torch.from_numpy(g["linear_expected"]), atol=1e-4)
        c1 = nn.Conv1d(3, 3, 1, bias=False); c1.weight.data.copy_(w.view(3, 3, 1))
        out = bnn.prepare_binary_model(c1, CFG)(torch.from_numpy(g["data"])[:, :, :, 0].reshape(1, 3, 2))
        assert torch.allclose(out, torch.from_numpy(g["conv1d_expected"]), atol=1e-4)
        c2 = nn.Conv2d(3, 3, 1, bias=False); c2.weight.data.copy_(w.view(3, 3, 1, 1))
        out = bnn.prepare_binary_model(c2, CFG)(torch.from_numpy(g["data"]))
        assert torch.allclose(out, torch.from_numpy(g["conv2d_expected"]), atol=1e-4)
        assert torch.equal(BasicInputBinarizer()(torch.from_numpy(g["sign_in"][:5])), torch.from_numpy(g["sign_ref"][:5]))


def test_state_dict_round_trip_is_bit_exact():
    # reference test/test_binarize.py:95-110
    torch.manual_seed(3)
    base, x = small_net(), torch.rand(1, 3, 8, 8)
    with runtime.floatsim_enabled(), torch.no_grad():
        model = bnn.prepare_binary_model(copy.deepcopy(base), CFG).eval()
        out1 = model(x.clone())
        state = model.state_dict()
        fresh = copy.deepcopy(base)
        for m in fresh.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                m.reset_parameters()
        fresh = bnn.prepare_binary_model(fresh, CFG).eval()
        fresh.load_state_dict(state)
        assert torch.equal(out1, fresh(x.clone()))


def test_ste_backward_and_other_binarizers():
    x = torch.tensor([0.5, -0.2, 1.5, -3.0], requires_grad=True)
    SignActivation.apply(x).sum().backward()
    assert torch.equal(x.grad, torch.tensor([1.0, 1.0, 0.0, 0.0]))
    assert torch.equal(AdvancedInputBinarizer()(torch.tensor([0.1, -0.1, 0.0])), torch.tensor([1.0, -1.0, 0.0]))
    y = StochasticInputBinarizer()(torch.zeros(1000))
    assert set(y.unique().tolist()) <= {-1.0, 1.0}
    conv = nn.Conv2d(4, 4, 3, padding=1)
    k = XNORScaleBinarizer(conv)
    assert k(torch.ones(1, 4, 5, 5), torch.ones(1, 4, 5, 5))[0, 0, 2, 2] == 1.0
    with pytest.raises(Exception):
        BasicScaleBinarizer(nn.ReLU())


def test_update_copies_learned_scales():
    a = Conv2d.from_module(nn.Conv2d(3, 4, 1), CFG)
    a.activation_post_process.alpha.data.fill_(2.5)
    b = Conv2d.from_module(a, CFG, update=True)
    assert torch.equal(b.activation_post_process.alpha, a.activation_post_process.alpha)
    c = Conv2d.from_module(a, CFG)
    assert torch.equal(c.activation_post_process.alpha, torch.ones(1, 4, 1, 1))


def test_lowering_analysis():
    from bnn_b200.layers.binary_modules import NotLowerable
    ok = Conv2d.from_module(nn.Conv2d(64, 64, 3, padding="same"), CFG)
    low = ok._lowering()
    assert low.fused_post and low.compute_alpha and not low.center_weights
    assert ok._resolved_padding() == (1, 1)
    assert Conv2d.from_module(nn.Conv2d(64, 64, 3, groups=2), CFG)._lowering().compute_alpha     # grouped: per-group launches
    with pytest.raises(NotLowerable):
        Conv1d.from_module(nn.Conv1d(64, 64, 3, groups=2), CFG)._lowering()
    with pytest.raises(NotLowerable):
        Conv2d.from_module(nn.Conv2d(64, 64, 3, padding=1, padding_mode="reflect"), CFG)._lowering()
    sto = bnn.BConfig(StochasticInputBinarizer, bnn.Identity, XNORWeightBinarizer)
    with pytest.raises(NotLowerable):
        Conv2d.from_module(nn.Conv2d(64, 64, 3), sto)._lowering()
    odd = bnn.BConfig(BasicInputBinarizer, BasicScaleBinarizer.with_args(shape=[1, 1, 1, 1]), XNORWeightBinarizer)
    assert Conv2d.from_module(nn.Conv2d(64, 64, 3), odd)._lowering().fused_post is False


@pytest.mark.skipif(not os.path.isdir("/root/reference/bnn"), reason="reference tree only exists in the build container")
def test_plugs_into_the_reference_prepare_binary_model():
    """Drop-in mode: the REFERENCE's prepare_binary_model with our mapping installs the CUDA layers,
    and a reference BConfig (reference binarizer classes) lowers."""
    import importlib
    import importlib.util
    import sys
    if "bnn_ref" not in sys.modules:
        spec = importlib.util.spec_from_file_location("bnn_ref", "/root/reference/bnn/__init__.py",
                                                      submodule_search_locations=["/root/reference/bnn"])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["bnn_ref"] = mod
        sys.dont_write_bytecode = True
        spec.loader.exec_module(mod)
    ref = sys.modules["bnn_ref"]
    rops = importlib.import_module("bnn_ref.ops")
    rcfg = ref.BConfig(activation_pre_process=rops.BasicInputBinarizer, activation_post_process=rops.BasicScaleBinarizer,
                       weight_pre_process=rops.XNORWeightBinarizer.with_args(center_weights=True))
    mapping = bnn.mapping_for_reference(ref)
    model = ref.prepare_binary_model(small_net(), rcfg, modules_mapping=mapping, ignore_layers_name=["8"])
    assert type(model[0]) is Conv2d and type(model[3]) is Conv2d and type(model[8]) is nn.Linear
    low = model[3]._lowering()
    assert low.center_weights and low.fused_post
    # the reference's own ResNet object, converted through the reference's entry point, is fusable as a whole
    from bnn_b200 import fuse
    rr = importlib.import_module("bnn_ref.models.resnet")
    r18 = ref.prepare_binary_model(rr.resnet18(), rcfg, modules_mapping=mapping, ignore_layers_name=["_first_", "_last_"])
    eng = fuse.optimize(r18.eval())
    assert isinstance(eng, fuse.FusedResNet) and eng.fused_blocks == 8 and eng.stem.ok
    # a model already converted by the reference converts too
    refmodel = ref.prepare_binary_model(small_net(), rcfg)
    model2 = ref.prepare_binary_model(refmodel, rcfg, modules_mapping=mapping)
    assert type(model2[0]) is Conv2d and type(model2[8]) is Linear


def test_binary_chef_runs_staged_recipes():
    """reference test/test_engine.py:39-66 re-stated: step count, per-step types, learned scales unchanged
    (fresh ones, like upstream) across re-conversion."""
    from bnn_b200.engine import BinaryChef
    from conftest import ROOT
    chef = BinaryChef(os.path.join(ROOT, "tests", "assets", "recipe.yaml"))
    assert len(chef) == chef.get_num_steps() == 3
    model = small_net()
    model = chef.next(model)
    assert type(model[0]) is nn.Conv2d and hasattr(model[3], "bconfig")
    assert isinstance(model[3].weight_pre_process, nn.Identity)
    model = chef.next(model)
    assert isinstance(model[3].weight_pre_process, XNORWeightBinarizer) and model[3].weight_pre_process.center_weights
    alpha = model[3].activation_post_process.alpha
    model = chef.next(model)
    assert isinstance(model[0], Conv2d) and isinstance(model[8], Linear)         # nothing ignored in the last stage
    assert isinstance(model[3].activation_pre_process, AdvancedInputBinarizer) and model[3].activation_pre_process.t == 3
    assert model[3].activation_pre_process.derivative_funct is torch.tanh
    assert not model[3].weight_pre_process.center_weights
    assert torch.equal(alpha, torch.ones_like(alpha))
    with pytest.raises(AssertionError):
        chef.next(model)
