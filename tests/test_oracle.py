"""Pins the oracle (oracle/) against the reference: its own known-answer vectors and outputs of the
real reference generated in the build container (tests/golden/*.npz).  CPU only."""
import numpy as np
import pytest
import torch

import cases
from conftest import rel_err
from oracle import c_oracle as co
from oracle import floatsim as fs


def _t(a):
    return None if a is None else torch.from_numpy(a)


def test_unit_vectors_reference_known_answers(golden_units):
    g = golden_units
    data, w = g["data"], g["weights"]
    ones3 = np.ones(3, np.float32)
    # Linear(3,3): reference test/test_layers.py:30-38
    y = fs.linear(_t(data[:, :, 0, 0].reshape(1, 3)), _t(w.reshape(3, 3)), None, _t(ones3)).numpy()
    assert np.allclose(y, g["linear_expected"], atol=1e-4)
    assert np.array_equal(y, g["linear_ref"])
    yc = co.floatsim_linear(data[:, :, 0, 0].reshape(1, 3), w.reshape(3, 3), None, ones3, False, True)
    assert np.allclose(yc, g["linear_expected"], atol=1e-4)
    # Conv1d(3,3,1): test/test_layers.py:40-50
    y = fs.conv1d(_t(np.ascontiguousarray(data[:, :, :, 0].reshape(1, 3, 2))), _t(w.reshape(3, 3, 1)), None, _t(ones3)).numpy()
    assert np.allclose(y, g["conv1d_expected"], atol=1e-4)
    assert np.array_equal(y, g["conv1d_ref"])
    # Conv2d(3,3,1): test/test_layers.py:52-67
    y = fs.conv2d(_t(data), _t(w.reshape(3, 3, 1, 1)), None, _t(ones3)).numpy()
    assert np.allclose(y, g["conv2d_expected"], atol=1e-4)
    assert np.array_equal(y, g["conv2d_ref"])
    geo = co.geom(1, 3, 2, 2, 3, 1, 1)
    yc = co.floatsim_conv2d(data, w.reshape(3, 3, 1, 1), None, ones3, geo, False, True)
    assert np.allclose(yc, g["conv2d_expected"], atol=1e-4)
    ab = co.pack_act(data)
    wb, alpha, nz = co.pack_weight(w.reshape(3, 3, 1, 1), False, True)
    assert nz == 0
    yi = co.bconv2d(ab, wb, alpha, None, ones3, geo)
    assert np.allclose(yi, g["conv2d_expected"], atol=1e-4)


def test_sign_edge_cases(golden_units):
    # test/test_binarize.py:118-120 plus +-0, denormal-ish and NaN
    x, ref = golden_units["sign_in"], golden_units["sign_ref"]
    assert np.array_equal(torch.sign(_t(x)).numpy(), ref, equal_nan=True)
    ab = co.pack_act(x.reshape(1, -1, 1, 1))
    s, m = int(ab[0, 0, 0, 0, 0]), int(ab[0, 0, 0, 0, 2])
    got = np.array([((s >> i) & 1) * 2 - 1 if (m >> i) & 1 else 0 for i in range(x.size)], np.float32)
    want = np.nan_to_num(ref, nan=0.0)   # the mask plane encodes sign(nan) as 0, like sign(+-0)
    assert np.array_equal(got, want)


def _run_floatsim_torch(case):
    x, w, bias, post = cases.make_inputs(case)
    hp = cases.hyper(case)
    if case["kind"] == "conv2d":
        return fs.conv2d(_t(x), _t(w), _t(bias), _t(post), hp["stride"], hp["pad"], hp["dil"], hp["alpha"], hp["center"],
                         hp["groups"]).numpy()
    if case["kind"] == "conv1d":
        return fs.conv1d(_t(x), _t(w), _t(bias), _t(post), hp["stride"], hp["pad"], hp["dil"], hp["alpha"], hp["center"]).numpy()
    return fs.linear(_t(x), _t(w), _t(bias), _t(post), hp["alpha"], hp["center"]).numpy()


def hp_groups(case):
    return int(case.get("groups", 1))


def _as_conv2d(case):
    """(x4d, w4d, geom, unflatten) so that conv1d / linear go through the conv2d C entry points."""
    x, w, bias, post = cases.make_inputs(case)
    hp = cases.hyper(case)
    if case["kind"] == "conv2d":
        n, c, h, wd = x.shape
        g = co.geom(n, c, h, wd, w.shape[0], w.shape[2], w.shape[3], hp["stride"], hp["pad"], hp["dil"])
        return x, w, bias, post, g, hp, lambda y: y
    if case["kind"] == "conv1d":
        n, c, l = x.shape
        g = co.geom(n, c, 1, l, w.shape[0], 1, w.shape[2], (1, hp["stride"][0]), (0, hp["pad"][0]), (1, hp["dil"][0]))
        return x[:, :, None, :], w[:, :, None, :], bias, post, g, hp, lambda y: y[:, :, 0, :]
    rows = x.reshape(-1, x.shape[-1])
    x4 = np.ascontiguousarray(rows.T)[None, :, None, :]            # [1, in, 1, rows]
    g = co.geom(1, rows.shape[1], 1, rows.shape[0], w.shape[0], 1, 1)
    lead = x.shape[:-1]
    return x4, w[:, :, None, None], bias, post, g, hp, lambda y: np.ascontiguousarray(y[0, :, 0, :].T).reshape(*lead, -1)


@pytest.mark.parametrize("case", cases.CASES, ids=[c["name"] for c in cases.CASES])
def test_layer_cases_against_reference_outputs(case, golden_layers):
    ref = golden_layers[case["name"]]
    # (1) torch restatement: same torch calls as the reference => tight
    y = _run_floatsim_torch(case)
    assert y.shape == ref.shape
    assert rel_err(y, ref) <= 1e-6, rel_err(y, ref)
    if hp_groups(case) > 1:
        # the C restatement has no groups argument: run it per group on channel slices, like the product does
        x, w, bias, post = cases.make_inputs(case)
        hp = cases.hyper(case)
        G, cin_g, cout_g = hp["groups"], w.shape[1], w.shape[0] // hp["groups"]
        outs = []
        for i in range(G):
            xs = np.ascontiguousarray(x[:, i * cin_g:(i + 1) * cin_g])
            ws = np.ascontiguousarray(w[i * cout_g:(i + 1) * cout_g])
            geo = co.geom(xs.shape[0], cin_g, xs.shape[2], xs.shape[3], cout_g, w.shape[2], w.shape[3], hp["stride"], hp["pad"], hp["dil"])
            sl = slice(i * cout_g, (i + 1) * cout_g)
            wb, alpha, nz = co.pack_weight(ws, hp["center"], hp["alpha"])
            assert nz == 0
            outs.append(co.bconv2d(co.pack_act(xs), wb, alpha if hp["alpha"] else None, None if bias is None else bias[sl],
                                   None if post is None else post[sl], geo))
        assert rel_err(np.concatenate(outs, 1), ref) <= 1e-5
        return
    # (2) plain-C float restatement and (3) packed integer formulation
    x4, w4, bias, post, g, hp, unflat = _as_conv2d(case)
    yc = unflat(co.floatsim_conv2d(x4, w4, bias, post, g, hp["center"], hp["alpha"]))
    assert rel_err(yc, ref) <= 1e-5, rel_err(yc, ref)
    ab = co.pack_act(x4)
    wb, alpha, nz = co.pack_weight(w4, hp["center"], hp["alpha"])
    assert nz == 0
    yi = unflat(co.bconv2d(ab, wb, alpha if hp["alpha"] else None, bias, post, g))
    assert rel_err(yi, ref) <= 1e-5, rel_err(yi, ref)


def test_integer_dot_properties():
    case = cases.by_name("ragged_c70_s21")
    x4, w4, _, _, g, hp, _ = _as_conv2d(case)
    ab = co.pack_act(x4)
    wb, _, _ = co.pack_weight(w4, hp["center"], hp["alpha"])
    dot = co.bconv2d_dot(ab, wb, g)
    abn = co.pack_act(-x4)
    assert np.array_equal(ab[..., 2:], abn[..., 2:])                     # mask planes unchanged
    assert np.array_equal(co.bconv2d_dot(abn, wb, g), -dot)       # antisymmetry in x
    k = 70 * 9
    assert np.abs(dot).max() <= k
    # brute force on the float side: dot == conv(sign x, sign wc) exactly
    t = torch.nn.functional.conv2d(torch.sign(torch.from_numpy(x4)).double(),
                                   torch.sign(torch.from_numpy(w4 - w4.mean(1, keepdims=True))).double(),
                                   None, hp["stride"], hp["pad"], hp["dil"]).numpy()
    assert np.array_equal(dot, t.astype(np.int32))


def test_zero_weights_are_reported():
    w = np.random.default_rng(0).standard_normal((4, 8, 1, 1)).astype(np.float32)
    w[1, 3] = 0.0
    w[2, 5] = -0.0
    _, _, nz = co.pack_weight(w, False, True)
    assert nz == 2


def test_whole_model_twin_matches_reference_logits(golden_models):
    """bnn_b200-prepared workload -> float-simulated twin (oracle) reproduces the REAL reference's logits
    on the seeded model, and the seeded parameters are the reference's (checksum)."""
    import torch.nn as nn
    import bnn_b200 as bnn
    from bnn_b200 import workloads
    from bnn_b200.ops import BasicInputBinarizer, XNORWeightBinarizer
    torch.set_grad_enabled(False)
    for variant in ("basic_relu", "pre_prelu"):
        torch.manual_seed(0)
        m = workloads.resnet18() if variant == "basic_relu" else workloads.resnet18(workloads.PreBasicBlock, nn.PReLU)
        cfg = bnn.BConfig(BasicInputBinarizer, bnn.Identity,
                          XNORWeightBinarizer.with_args(compute_alpha=True, center_weights=True))
        m = bnn.prepare_binary_model(m, cfg, ignore_layers_name=["_first_", "_last_"])
        assert sum(isinstance(x, bnn.layers.Conv2d) for x in m.modules()) == 19
        workloads.randomize_batchnorm(m, seed=1)
        chk = np.array([float(p.double().abs().sum()) for p in m.state_dict().values() if p.dtype.is_floating_point])
        assert np.allclose(chk, golden_models[variant + "_checksum"], rtol=1e-12)
        twin = fs.mirror_model(m)
        x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(0))
        y = twin(x).numpy()
        assert rel_err(y, golden_models[variant + "_logits"]) <= 1e-6
    torch.set_grad_enabled(True)


@pytest.mark.parametrize("which,nconv", [("resnet50", 52), ("hblock", 16)])
def test_config34_twin_matches_reference_logits(which, nconv, golden_cfg34):
    """BASELINE configs[2] / [3]: the seeded workload definitions reproduce the real reference's parameters (checksum)
    and the oracle's float-simulated twin reproduces the real reference's logits (tests/golden/models_cfg34.npz)."""
    from conftest import build_config34
    import bnn_b200 as bnn
    torch.set_grad_enabled(False)
    try:
        m, x = build_config34(which)
        assert sum(isinstance(mod, bnn.layers.Conv2d) for mod in m.modules()) == nconv
        chk = np.array([float(p.double().abs().sum()) for p in m.state_dict().values() if p.dtype.is_floating_point])
        assert np.allclose(chk, golden_cfg34[which + "_checksum"], rtol=1e-12)
        y = fs.mirror_model(m)(x).numpy()
        assert rel_err(y, golden_cfg34[which + "_logits"]) <= 1e-6
    finally:
        torch.set_grad_enabled(True)


# ----------------------------------------------------------------------------------------------
# cross-module fusion epilogue (struct bnn_epilogue) and the pooled / affine bit-pack
# ----------------------------------------------------------------------------------------------
FUSED_CASES = [
    dict(name="basic_mid", cin=64, cout=64, hw=(9, 11), bn=True, act=1, bits=True, out=False),
    dict(name="basic_out", cin=64, cout=64, hw=(9, 11), bn=True, act=1, res="pre", bits=True, out=True),
    dict(name="pre_mid", cin=128, cout=128, hw=(7, 7), act=2, bits=True, out=False, nx=True),
    dict(name="pre_out", cin=128, cout=128, hw=(7, 7), act=2, res="post", bits=True, out=True, nx=True),
    dict(name="shortcut", cin=64, cout=128, hw=(8, 8), k=1, pad=0, bn=True, out=True, bits=False),
    dict(name="ragged_c1", cin=70, cout=24, hw=(6, 5), bn=True, act=1, res="pre", bits=True, out=True, bias=True, post=True),
    dict(name="c96_s2", cin=64, cout=96, hw=(10, 10), stride=2, bn=True, act=2, bits=True, out=True, nx=True),
    dict(name="hblock_stage", cin=128, cout=64, hw=(8, 8), res="post", bits=True, out=True, nx=True, nx_relu=True, bits_pre=True),
    # complete pixel groups for every P in {8, 7, 4} and complete channel blocks: the lean NHWC epilogue (EPI 3) runs
    dict(name="lean_mid", cin=64, cout=128, hw=(3, 56), bn=True, act=1, bits=True, out=False),
    dict(name="lean_out", cin=64, cout=128, hw=(3, 56), bn=True, act=1, res="pre", bits=True, out=True),
    dict(name="lean_s2", cin=128, cout=128, hw=(6, 112), stride=2, bn=True, act=1, bits=True, out=True),
    dict(name="lean_out_only", cin=64, cout=256, hw=(2, 56), bn=True, act=1, res="pre", bits=False, out=True, bias=True, post=True),
    # ... and its general form: pre-activation blocks (PReLU, shortcut after the activation, next BatchNorm before the sign)
    dict(name="lean_pre_mid", cin=64, cout=128, hw=(3, 56), act=2, bits=True, out=False, nx=True),
    dict(name="lean_pre_out", cin=128, cout=128, hw=(3, 56), act=2, res="post", bits=True, out=True, nx=True),
    dict(name="lean_res_after_noact", cin=64, cout=256, hw=(2, 56), bn=True, res="post", bits=True, out=True, nx=True, post=True),
    # 1x1 kernels over several 64-channel chunks: the chunk-triple carry-save path (5 chunks = one triple + two singles)
    dict(name="k1_c320", cin=320, cout=96, hw=(6, 10), k=1, pad=0, bn=True, act=1, bits=True, out=True),
    dict(name="k1_c192_lean", cin=192, cout=128, hw=(2, 56), k=1, pad=0, bn=True, act=1, res="pre", bits=True, out=True),
    dict(name="k1_c512_s2", cin=512, cout=64, hw=(8, 8), k=1, pad=0, stride=2, bn=True, act=2, bits=True, out=True, nx=True),
]


def make_fused_inputs(fc):
    rng = np.random.default_rng(4242 + [c["name"] for c in FUSED_CASES].index(fc["name"]))
    k, pad, stride = fc.get("k", 3), fc.get("pad", 1), fc.get("stride", 1)
    h, w = fc["hw"]
    n = 2
    x = np.maximum(rng.standard_normal((n, fc["cin"], h, w)), 0).astype(np.float32)
    wt = (rng.standard_normal((fc["cout"], fc["cin"], k, k)) * 0.05).astype(np.float32)
    g = co.geom(n, fc["cin"], h, w, fc["cout"], k, k, (stride, stride), (pad, pad), (1, 1))
    ho, wo = co.out_hw(g)
    co_ = fc["cout"]
    d = dict(x=x, w=wt, g=g,
             bias=(rng.standard_normal(co_) * 0.3).astype(np.float32) if fc.get("bias") else None,
             post=(0.5 + rng.random(co_)).astype(np.float32) if fc.get("post") else None,
             bn=((0.5 + rng.random(co_)).astype(np.float32) * 3, (rng.standard_normal(co_) * 0.3).astype(np.float32)) if fc.get("bn") else None,
             residual=rng.standard_normal((n, co_, ho, wo)).astype(np.float32) if fc.get("res") else None,
             res_after=fc.get("res") == "post", act=fc.get("act", 0),
             slope=(rng.random(co_) * 0.5).astype(np.float32) if fc.get("act", 0) == 2 else None,
             nx=((0.5 + rng.random(co_)).astype(np.float32), (rng.standard_normal(co_) * 0.2).astype(np.float32)) if fc.get("nx") else None,
             want_out=fc.get("out", True), want_bits=fc.get("bits", False), nx_relu=bool(fc.get("nx_relu")),
             bits_pre=bool(fc.get("bits_pre")))
    return d


@pytest.mark.parametrize("fc", FUSED_CASES, ids=[c["name"] for c in FUSED_CASES])
def test_fused_epilogue_oracle_against_torch_composition(fc):
    """oracle (integer path + fused epilogue) == the module sequence the reference's blocks execute:
    conv (float-sim) -> BatchNorm(eval, folded) -> (+res) -> ReLU/PReLU -> (+res); bits == sign of the result."""
    d = make_fused_inputs(fc)
    g = d["g"]
    ab = co.pack_act(d["x"])
    wb, alpha, nz = co.pack_weight(d["w"], True, True)
    assert nz == 0
    out, bits = co.bconv2d_fused(ab, wb, g, scale=alpha, bias=d["bias"], post=d["post"], bn=d["bn"], residual=d["residual"],
                                 residual_after_act=d["res_after"], act=d["act"], act_slope=d["slope"], want_out=True,
                                 want_bits=True, nx=d["nx"], nx_relu=d["nx_relu"], bits_before_residual=d["bits_pre"])
    y = fs.conv2d(_t(d["x"]), _t(d["w"]), _t(d["bias"]), _t(d["post"]), (g.stride_h, g.stride_w), (g.pad_h, g.pad_w),
                  (1, 1), True, True)
    v = lambda a: _t(a).view(1, -1, 1, 1)
    if d["bn"] is not None:
        y = y * v(d["bn"][0]) + v(d["bn"][1])
    if d["residual"] is not None and not d["res_after"]:
        y = y + _t(d["residual"])
    if d["act"] == 1:
        y = torch.relu(y)
    elif d["act"] == 2:
        y = torch.nn.functional.prelu(y, _t(d["slope"]))
    y_before = y
    if d["residual"] is not None and d["res_after"]:
        y = y + _t(d["residual"])
    assert rel_err(out, y.numpy()) <= 1e-5
    if d["bits_pre"]:
        # HBlock stage: the next conv sees relu(bn(conv)); the block output is conv + shortcut
        nb = torch.relu(y_before * v(d["nx"][0]) + v(d["nx"][1])).numpy()
        assert (bits != co.pack_act(nb)).sum() <= 2          # fma vs mul+add at the knife edge
        return
    # emitted planes are the bit-pack of the oracle's own fp32 result (with the next layer's affine, which the
    # epilogue applies as one fma while the stand-alone pack rounds the product first: identical except when
    # the affine lands within an ulp of zero)
    want_bits = co.pack_act(out, pre_scale=None if d["nx"] is None else d["nx"][0],
                            pre_shift=None if d["nx"] is None else d["nx"][1])
    assert (bits != want_bits).sum() <= (1 if d["nx"] is not None else 0)


def test_avgpool_pack_matches_torch():
    rng = np.random.default_rng(7)
    for shape, k, ceil in (((2, 64, 8, 8), 2, True), ((1, 70, 7, 9), 2, True), ((1, 64, 7, 9), 2, False), ((1, 3, 9, 9), 3, True)):
        x = np.maximum(rng.standard_normal(shape), 0).astype(np.float32)
        pooled = torch.nn.functional.avg_pool2d(_t(x), k, k, 0, ceil_mode=ceil, count_include_pad=False).numpy()
        assert np.array_equal(co.pack_act(x, pool=k, ceil_mode=ceil), co.pack_act(pooled))
    x = rng.standard_normal((2, 64, 5, 5)).astype(np.float32)
    s, h = (0.5 + rng.random(64)).astype(np.float32), rng.standard_normal(64).astype(np.float32)
    assert np.array_equal(co.pack_act(x, pre_scale=s, pre_shift=h), co.pack_act(x * s[None, :, None, None] + h[None, :, None, None]))


def test_stem_oracle_matches_torch_modules():
    """orc_stem == conv7x7/2 -> BatchNorm(eval) -> ReLU -> MaxPool(3,2,1) as torch executes the reference's stem."""
    rng = np.random.default_rng(11)
    for hw in ((64, 64), (37, 52)):
        x = rng.standard_normal((2, 3) + hw).astype(np.float32)
        w = (rng.standard_normal((64, 3, 7, 7)) * 0.1).astype(np.float32)
        g, h = (0.5 + rng.random(64)).astype(np.float32), (rng.standard_normal(64) * 0.3).astype(np.float32)
        out, bits = co.stem(x, w, g, h)
        y = torch.nn.functional.conv2d(_t(x), _t(w), None, 2, 3)
        y = torch.relu(y * _t(g).view(1, -1, 1, 1) + _t(h).view(1, -1, 1, 1))
        y = torch.nn.functional.max_pool2d(y, 3, 2, 1).permute(0, 2, 3, 1).numpy()
        assert out.shape == y.shape
        assert rel_err(out, y) <= 1e-5
        assert np.array_equal(bits, co.pack_act(np.ascontiguousarray(out.transpose(0, 3, 1, 2))))
