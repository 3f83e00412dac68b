"""Parity tests proper: the CUDA path (through the C ABI) against the oracle and the committed
reference outputs.  Integer work is compared bit-exactly; fp32 outputs within 1e-5 of max|ref|
(north_star tolerance is 1e-3).  Run on the B200 box:  pytest tests -m gpu."""
import numpy as np
import pytest
import torch

import cases
from conftest import rel_err
from oracle import c_oracle as co
from test_oracle import _as_conv2d

import bnn_b200 as bnn
from bnn_b200 import functional as BF
from bnn_b200 import native, runtime

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def _bits(t):
    return t.cpu().numpy().view(np.uint32)


def test_library_reports_device():
    assert native.query(native.Q_DEVICE_SMS) > 0
    assert torch.cuda.get_device_capability(0)[0] == 10


@pytest.mark.parametrize("shape,layout", [((2, 64, 5, 7), "nchw"), ((1, 3, 4, 4), "nchw"), ((2, 70, 6, 7), "nchw"),
                                          ((2, 128, 9, 11), "nhwc"), ((3, 200, 3, 5), "sliced"), ((1, 512, 7, 7), "nchw"),
                                          ((4, 64, 56, 56), "nchw")])
def test_pack_activations_bit_exact(shape, layout):
    rng = np.random.default_rng(5)
    x = rng.standard_normal(shape).astype(np.float32)
    x[rng.random(shape) < 0.4] = 0.0
    flat = x.reshape(-1)
    flat[::13] = -0.0
    flat[5::97] = np.nan
    flat[7::101] = 1e-42        # denormal, still > 0
    xt = torch.from_numpy(x).to(DEV)
    if layout == "nhwc":
        xt = xt.contiguous(memory_format=torch.channels_last)
    elif layout == "sliced":
        big = torch.zeros((shape[0], shape[1] + 3, shape[2], shape[3] + 2), device=DEV)
        big[:, 1:-2, :, 1:-1] = xt
        xt = big[:, 1:-2, :, 1:-1]
        assert not xt.is_contiguous()
    packed = BF.pack_activations(xt)
    ab = co.pack_act(x)
    assert np.array_equal(_bits(packed.bits), ab)


@pytest.mark.parametrize("shape", [(3, 128, 5, 7), (1, 256, 3, 3), (2, 512, 2, 3), (5, 64, 3, 3), (2, 256, 16, 16)])
@pytest.mark.parametrize("pre,pre_relu", [(False, False), (True, False), (True, True)])
def test_pack_dense_nhwc_fast_path_bit_exact(shape, pre, pre_relu):
    """Dense channels_last tensors with C in {64,128,256,512} take the streaming kernel (4 pixels per warp): ragged pixel
    counts, NaN / -0 / denormals, the pre-sign affine and the ReLU-in-front form, all against the oracle."""
    rng = np.random.default_rng(8)
    x = rng.standard_normal(shape).astype(np.float32)
    x[rng.random(shape) < 0.3] = 0.0
    flat = x.reshape(-1)
    flat[::13] = -0.0
    flat[5::97] = np.nan
    flat[7::101] = 1e-42
    c = shape[1]
    sc = (0.5 + rng.random(c)).astype(np.float32) if pre else None
    sh = rng.standard_normal(c).astype(np.float32) if pre else None
    xt = torch.from_numpy(x).to(DEV).contiguous(memory_format=torch.channels_last)
    got = BF.pack_activations(xt, pre=None if not pre else (torch.from_numpy(sc).to(DEV), torch.from_numpy(sh).to(DEV)),
                              pre_relu=pre_relu)
    want = co.pack_act(x, pre_scale=sc, pre_shift=sh, pre_relu=pre_relu)
    assert np.array_equal(_bits(got.bits), want)


@pytest.mark.parametrize("wshape,center,alpha", [((64, 64, 3, 3), True, True), ((40, 70, 3, 3), True, True),
                                                 ((3, 3, 1, 1), False, True), ((128, 64, 1, 1), False, False),
                                                 ((1000, 512), True, True), ((8, 16, 5), False, True),
                                                 ((512, 512, 3, 3), True, True)])
def test_pack_weights_bits_exact_alpha_close(wshape, center, alpha):
    rng = np.random.default_rng(11)
    w = (rng.standard_normal(wshape) * 0.05).astype(np.float32)
    packed = BF.pack_weights(torch.from_numpy(w).to(DEV), center, alpha)
    wb, al, nz = co.pack_weight(w, center, alpha)
    assert packed.n_zero == nz == 0
    assert np.array_equal(_bits(packed.bits), wb)
    assert rel_err(packed.alpha.cpu().numpy(), al) <= 1e-6


def test_pack_weights_counts_exact_zeros():
    w = torch.randn(4, 8, 1, 1, device=DEV)
    w[1, 3] = 0.0
    w[2, 5] = -0.0
    assert BF.pack_weights(w, False, True).n_zero == 2


FLAG_SETS = [native.F_STAGE_LDG, 0, native.F_NO_CSA]


UNGROUPED = [c for c in cases.CASES if c.get("groups", 1) == 1]


@pytest.mark.parametrize("flags", FLAG_SETS, ids=["ldg", "tma", "nocsa"])
@pytest.mark.parametrize("case", UNGROUPED, ids=[c["name"] for c in UNGROUPED])
def test_conv_kernel_against_oracle_and_reference(case, flags, golden_layers):
    """C ABI level: packed conv == oracle integer path bit-exactly (dot), == reference within TOL."""
    x4, w4, bias, post, g, hp, unflat = _as_conv2d(case)
    xt = torch.from_numpy(np.ascontiguousarray(x4)).to(DEV)
    act = BF.pack_activations(xt)
    wts = BF.pack_weights(torch.from_numpy(np.ascontiguousarray(w4)).to(DEV), hp["center"], hp["alpha"])
    stride, pad, dil = (g.stride_h, g.stride_w), (g.pad_h, g.pad_w), (g.dil_h, g.dil_w)
    # integer dot: scale = None, no bias / post -> exact small integers in fp32
    dot = BF.bconv2d(act, wts, None, None, stride, pad, dil, use_alpha=False, flags=flags).cpu().numpy()
    ab = co.pack_act(x4)
    wb, _, _ = co.pack_weight(w4, hp["center"], hp["alpha"])
    want = co.bconv2d_dot(ab, wb, g)
    assert np.array_equal(dot.astype(np.int32), want) and np.array_equal(dot, want.astype(np.float32))
    # fused epilogue vs the real reference's output
    bt = None if bias is None else torch.from_numpy(bias).to(DEV)
    pt = None if post is None else torch.from_numpy(post).to(DEV)
    y = BF.bconv2d(act, wts, bt, pt, stride, pad, dil, use_alpha=hp["alpha"], flags=flags).cpu().numpy()
    assert rel_err(unflat(y), golden_layers[case["name"]]) <= TOL


@pytest.mark.parametrize("case", cases.CASES, ids=[c["name"] for c in cases.CASES])
def test_module_api_against_reference(case, golden_layers):
    """Plugin level: nn module -> prepare_binary_model -> forward on the GPU, as a user would."""
    import torch.nn as nn
    from bnn_b200.ops import BasicInputBinarizer, BasicScaleBinarizer, XNORWeightBinarizer
    x, w, bias, post = cases.make_inputs(case)
    hp = cases.hyper(case)
    cfg = bnn.BConfig(BasicInputBinarizer, BasicScaleBinarizer if post is not None else bnn.Identity,
                      XNORWeightBinarizer.with_args(compute_alpha=hp["alpha"], center_weights=hp["center"]))
    if case["kind"] == "conv2d":
        m = nn.Conv2d(w.shape[1] * hp["groups"], w.shape[0], w.shape[2:], stride=hp["stride"], padding=hp["pad"],
                      dilation=hp["dil"], groups=hp["groups"], bias=bias is not None)
    elif case["kind"] == "conv1d":
        m = nn.Conv1d(w.shape[1], w.shape[0], w.shape[2], stride=hp["stride"], padding=hp["pad"], dilation=hp["dil"],
                      bias=bias is not None)
    else:
        m = nn.Linear(w.shape[1], w.shape[0], bias=bias is not None)
    m.weight.data.copy_(torch.from_numpy(w))
    if bias is not None:
        m.bias.data.copy_(torch.from_numpy(bias))
    m = bnn.prepare_binary_model(m.to(DEV), cfg).eval()
    if post is not None:
        m.activation_post_process.alpha.data.copy_(torch.from_numpy(post).reshape(m.activation_post_process.alpha.shape))
    before = native.launch_count()
    with torch.no_grad():
        y = m(torch.from_numpy(x).to(DEV))
    assert native.launch_count() >= before + 2 * hp["groups"]   # pack + conv (per group) really ran in the native library
    assert rel_err(y.cpu().numpy(), golden_layers[case["name"]]) <= TOL


def test_reference_known_answer_vectors_on_gpu(golden_units):
    import torch.nn as nn
    from bnn_b200.ops import BasicInputBinarizer, BasicScaleBinarizer, XNORWeightBinarizer
    cfg = bnn.BConfig(BasicInputBinarizer, BasicScaleBinarizer, XNORWeightBinarizer)
    g = golden_units
    w = torch.from_numpy(g["weights"])
    data = torch.from_numpy(g["data"]).to(DEV)
    with torch.no_grad():
        lin = nn.Linear(3, 3, bias=False); lin.weight.data.copy_(w.view(3, 3))
        out = bnn.prepare_binary_model(lin.to(DEV), cfg)(data[:, :, 0, 0].reshape(1, 3))
        assert np.allclose(out.cpu().numpy(), g["linear_expected"], atol=1e-4)
        c1 = nn.Conv1d(3, 3, 1, bias=False); c1.weight.data.copy_(w.view(3, 3, 1))
        out = bnn.prepare_binary_model(c1.to(DEV), cfg)(data[:, :, :, 0].reshape(1, 3, 2))
        assert np.allclose(out.cpu().numpy(), g["conv1d_expected"], atol=1e-4)
        c2 = nn.Conv2d(3, 3, 1, bias=False); c2.weight.data.copy_(w.view(3, 3, 1, 1))
        out = bnn.prepare_binary_model(c2.to(DEV), cfg)(data)
        assert np.allclose(out.cpu().numpy(), g["conv2d_expected"], atol=1e-4)


def test_weight_cache_follows_load_state_dict_and_inplace_updates():
    import torch.nn as nn
    from bnn_b200.ops import BasicInputBinarizer, XNORWeightBinarizer
    cfg = bnn.BConfig(BasicInputBinarizer, bnn.Identity, XNORWeightBinarizer)
    torch.manual_seed(0)
    a = bnn.prepare_binary_model(nn.Conv2d(64, 64, 3, padding=1, bias=False).to(DEV), cfg).eval()
    b = bnn.prepare_binary_model(nn.Conv2d(64, 64, 3, padding=1, bias=False).to(DEV), cfg).eval()
    x = torch.randn(2, 64, 9, 9, device=DEV)
    with torch.no_grad():
        ya, yb = a(x), b(x)
        assert not torch.equal(ya, yb)
        b.load_state_dict(a.state_dict())
        assert torch.equal(a(x), b(x))
        a.weight.neg_()                                   # in-place update bumps the version counter
        assert torch.equal(a(x), -ya)                     # repacked: alpha unchanged, every dot negated


def _sign_dot_f64(x, w_centred, stride, pad, dil, groups=1):
    """exact integer dot of the reference's ternary operands: conv(sign(x), sign(w)) in float64"""
    return torch.nn.functional.conv2d(torch.sign(torch.from_numpy(x).double()), torch.sign(torch.from_numpy(w_centred).double()),
                                      None, stride, pad, dil, groups).numpy()


@pytest.mark.parametrize("center", [False, True])
def test_exact_zero_weights_run_the_ternary_path(center):
    """sign(0) = 0 for weights too (reference bnn/ops.py:66,136): tensors with exactly-zero (centred) weights run as two
    binary launches (zeros packed as -1 / +1) + an exact integer mean.  Integer dots bit-exact, layer output vs the
    float simulation of the reference."""
    import torch.nn as nn
    from oracle import floatsim as fs
    from bnn_b200.ops import BasicInputBinarizer, BasicScaleBinarizer, XNORWeightBinarizer
    rng = np.random.default_rng(21)
    x = rng.standard_normal((2, 70, 9, 11)).astype(np.float32)
    x[rng.random(x.shape) < 0.3] = 0.0
    w = (rng.standard_normal((40, 70, 3, 3)) * 0.05).astype(np.float32)
    if center:
        w[3, :, 1, 1] = 0.25                  # a constant column: centring makes all 70 of them exactly zero
        w[17, :, 0, 2] = -0.5
        # the pack kernel's centring: per-tap mean over c_in summed in fp64 and rounded once, fp32 subtraction
        wc = w - (w.astype(np.float64).sum(1, keepdims=True) / w.shape[1]).astype(np.float32)
    else:
        w[rng.random(w.shape) < 0.2] = 0.0    # pruned weights
        w[5] = 0.0                            # a dead output channel: alpha = 0, every dot = 0
        wc = w
    nz_want = int((wc == 0).sum())
    assert nz_want > 0
    wts = BF.pack_weights(torch.from_numpy(w).to(DEV), center, True)
    assert wts.n_zero == nz_want and wts.hi is not None
    act = BF.pack_activations(torch.from_numpy(x).to(DEV))
    dot = BF.bconv2d(act, wts, None, None, (2, 1), (1, 1), (1, 1), use_alpha=False).cpu().numpy()
    assert np.array_equal(dot.astype(np.float64), _sign_dot_f64(x, wc, (2, 1), (1, 1), (1, 1)))
    # module level: bias + learned post scale, against the torch float simulation of the reference's forward
    cfg = bnn.BConfig(BasicInputBinarizer, BasicScaleBinarizer, XNORWeightBinarizer.with_args(compute_alpha=True, center_weights=center))
    m = nn.Conv2d(70, 40, 3, stride=(2, 1), padding=1)
    m.weight.data.copy_(torch.from_numpy(w))
    m = bnn.prepare_binary_model(m.to(DEV), cfg).eval()
    m.activation_post_process.alpha.data.uniform_(0.5, 1.5)
    with torch.no_grad():
        y = m(torch.from_numpy(x).to(DEV)).cpu()
        want = fs.conv2d(torch.from_numpy(x), m.weight.cpu(), m.bias.cpu(), m.activation_post_process.alpha.cpu().reshape(-1),
                         (2, 1), 1, 1, True, center)
    assert rel_err(y.numpy(), want.numpy()) <= TOL
    # Linear and a grouped conv with zeros
    lin = nn.Linear(100, 24)
    lin.weight.data[::3, ::7] = 0.0
    lin = bnn.prepare_binary_model(lin.to(DEV), bnn.BConfig(BasicInputBinarizer, bnn.Identity, XNORWeightBinarizer)).eval()
    xl = torch.randn(5, 100)
    with torch.no_grad():
        assert rel_err(lin(xl.to(DEV)).cpu().numpy(), fs.linear(xl, lin.weight.cpu(), lin.bias.cpu()).numpy()) <= TOL
    gc = nn.Conv2d(64, 64, 3, padding=1, groups=2, bias=False)
    gc.weight.data[gc.weight.data.abs() < 0.02] = 0.0
    gc = bnn.prepare_binary_model(gc.to(DEV), bnn.BConfig(BasicInputBinarizer, bnn.Identity, XNORWeightBinarizer)).eval()
    xg = torch.randn(2, 64, 7, 7)
    with torch.no_grad():
        want = fs.conv2d(xg, gc.weight.cpu(), None, None, 1, 1, 1, True, False, groups=2)
        assert rel_err(gc(xg.to(DEV)).cpu().numpy(), want.numpy()) <= TOL


@pytest.mark.parametrize("kind,cin,cout,k,hw", [("linear", 25088, 96, 1, (1, 3)), ("conv", 6144, 32, 3, (5, 6)),
                                               ("conv", 16448, 64, 1, (4, 4)), ("linear", 16390, 40, 1, (1, 1))])
def test_split_k_layers_bit_exact_vs_oracle(kind, cin, cout, k, hw):
    """Reductions too large for one CTA's shared memory (Linear(25088, .): 392 chunks; 3x3 over 6144 channels; more than
    16384 channels) are contracted chunk range by chunk range (bnn_conv_split / bnn_bconv2d_partial_fwd /
    bnn_dot_finish_f32).  Integer dots bit-exact against the oracle, outputs against its float simulation."""
    rng = np.random.default_rng(31)
    h, w_ = hw
    pad = k // 2
    x = np.maximum(rng.standard_normal((1, cin, h, w_)), -0.3).astype(np.float32)
    x[rng.random(x.shape) < 0.2] = 0.0
    wt = (rng.standard_normal((cout, cin, k, k)) * 0.05).astype(np.float32)
    bias = (rng.standard_normal(cout) * 0.3).astype(np.float32)
    post = (0.5 + rng.random(cout)).astype(np.float32)
    g = co.geom(1, cin, h, w_, cout, k, k, (1, 1), (pad, pad), (1, 1))
    ng = native.ConvGeom(1, cin, h, w_, cout, k, k, 1, 1, pad, pad, 1, 1)
    assert native.conv_split(ng)[1] > 1
    wb, alpha, nz = co.pack_weight(wt, True, True)
    assert nz == 0
    want_dot = co.bconv2d_dot(co.pack_act(x), wb, g)
    wts = BF.pack_weights(torch.from_numpy(wt).to(DEV), True, True)
    assert np.array_equal(wts.bits.cpu().numpy().view(np.uint32), wb)
    if kind == "linear":
        rows = x.reshape(cin, w_).T.copy()                        # [rows, features]
        act = BF.pack_activations(torch.from_numpy(rows).to(DEV), linear_rows=True)
        dot = BF.blinear(act, wts, None, None, use_alpha=False).cpu().numpy()          # [rows, out]
        assert np.array_equal(dot.T.reshape(want_dot.shape).astype(np.int32), want_dot)
        y = BF.blinear(act, wts, torch.from_numpy(bias).to(DEV), torch.from_numpy(post).to(DEV)).cpu().numpy()
        want = co.floatsim_linear(rows, wt.reshape(cout, cin), bias, post, True, True)
        assert rel_err(y, want) <= TOL
    else:
        act = BF.pack_activations(torch.from_numpy(x).to(DEV))
        dot = BF.bconv2d(act, wts, None, None, (1, 1), (pad, pad), (1, 1), use_alpha=False).cpu().numpy()
        assert np.array_equal(dot.astype(np.int32), want_dot)
        y = BF.bconv2d(act, wts, torch.from_numpy(bias).to(DEV), torch.from_numpy(post).to(DEV), (1, 1), (pad, pad), (1, 1)).cpu().numpy()
        want = co.floatsim_conv2d(x, wt, bias, post, g, True, True)
        assert rel_err(y, want) <= TOL


def test_large_linear_module_runs():
    """reference bnn/layers/linear.py:22-27 has no size limit: Linear(25088, 4096) (the VGG classifier shape) through the
    module API."""
    import torch.nn as nn
    from oracle import floatsim as fs
    from bnn_b200.ops import BasicInputBinarizer, XNORWeightBinarizer
    torch.manual_seed(3)
    lin = nn.Linear(25088, 4096)
    lin = bnn.prepare_binary_model(lin.to(DEV), bnn.BConfig(BasicInputBinarizer, bnn.Identity, XNORWeightBinarizer)).eval()
    x = torch.randn(4, 25088)
    with torch.no_grad():
        y = lin(x.to(DEV)).cpu()
        want = fs.linear(x, lin.weight.cpu(), lin.bias.cpu())
    assert rel_err(y.numpy(), want.numpy()) <= TOL


@pytest.mark.parametrize("cin,cout,k,s,pad,h,w", [(1153, 32, 5, 2, 2, 5, 57), (1153, 70, 5, 1, 2, 6, 300)])
def test_narrow_tile_fallback_geometries(cin, cout, k, s, pad, h, w):
    """19 input chunks x a 5x5 kernel: a tile spanning the whole output row does not fit shared memory, so the planner
    cuts the row into pieces.  Integer dots bit-exact against the oracle."""
    rng = np.random.default_rng(3)
    x = np.maximum(rng.standard_normal((1, cin, h, w)), 0).astype(np.float32)
    wt = (rng.standard_normal((cout, cin, k, k)) * 0.05).astype(np.float32)
    g = co.geom(1, cin, h, w, cout, k, k, (s, s), (pad, pad), (1, 1))
    wb, _, nz = co.pack_weight(wt, True, True)
    assert nz == 0
    want = co.bconv2d(co.pack_act(x), wb, None, None, None, g)
    plan = native.conv_plan(native.ConvGeom(1, cin, h, w, cout, k, k, s, s, pad, pad, 1, 1), 0, 148)
    assert plan["TW"] < want.shape[3]                       # the row really is cut
    act = BF.pack_activations(torch.from_numpy(x).to(DEV))
    wts = BF.pack_weights(torch.from_numpy(wt).to(DEV), True, True)
    got = BF.bconv2d(act, wts, None, None, (s, s), (pad, pad), (1, 1), use_alpha=False).cpu().numpy()
    assert np.array_equal(got, want)
