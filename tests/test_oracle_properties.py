"""Property tests of the oracle on random geometries (CPU only): the integer formulation on the packed layouts
(pack -> XNOR/AND popcount -> epilogue), the C float simulation and the torch float simulation that mirrors the
reference's own calls (bnn/layers/conv.py:90-97, bnn/ops.py:63-66,116-140,200-202) agree on every draw -- ragged
channel counts, strides, dilations, paddings larger than the kernel reach, zeros / -0 / tiny values in the input."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import c_oracle as co
from oracle import floatsim as fs


@st.composite
def geometries(draw):
    k = draw(st.sampled_from([1, 3, 5]))
    kw = draw(st.sampled_from([k, 1])) if k > 1 else 1
    dil = draw(st.sampled_from([1, 2]))
    stride = (draw(st.integers(1, 2)), draw(st.integers(1, 3)))
    pad = (draw(st.integers(0, 3)), draw(st.integers(0, 3)))
    cin, cout = draw(st.integers(1, 140)), draw(st.integers(1, 70))
    h = draw(st.integers(1, 9)) + dil * (k - 1)
    w = draw(st.integers(1, 9)) + dil * (kw - 1)
    return dict(n=draw(st.integers(1, 2)), cin=cin, cout=cout, h=h, w=w, kh=k, kw=kw, stride=stride, pad=pad, dil=(dil, dil),
                center=draw(st.booleans()), alpha=draw(st.booleans()), bias=draw(st.booleans()), post=draw(st.booleans()),
                zeros=draw(st.sampled_from([0.0, 0.3, 0.6])), seed=draw(st.integers(0, 2 ** 20)))


@settings(max_examples=60, deadline=None, derandomize=True)
@given(geometries())
def test_integer_path_equals_both_float_simulations(gm):
    rng = np.random.default_rng(gm["seed"])
    x = rng.standard_normal((gm["n"], gm["cin"], gm["h"], gm["w"])).astype(np.float32)
    x[rng.random(x.shape) < gm["zeros"]] = 0.0
    flat = x.reshape(-1)
    flat[::11] = -0.0
    flat[3::29] = 1e-42
    w = (rng.standard_normal((gm["cout"], gm["cin"], gm["kh"], gm["kw"])) * 0.1).astype(np.float32)
    bias = rng.standard_normal(gm["cout"]).astype(np.float32) if gm["bias"] else None
    post = (0.5 + rng.random(gm["cout"])).astype(np.float32) if gm["post"] else None
    g = co.geom(gm["n"], gm["cin"], gm["h"], gm["w"], gm["cout"], gm["kh"], gm["kw"], gm["stride"], gm["pad"], gm["dil"])
    wb, alpha, n_zero = co.pack_weight(w, gm["center"], gm["alpha"])
    if n_zero:                      # an exactly-zero centred weight has no 1-bit form: the product path refuses it too
        return
    got = co.bconv2d(co.pack_act(x), wb, alpha if gm["alpha"] else None, bias, post, g)
    sim_c = co.floatsim_conv2d(x, w, bias, post, g, gm["center"], gm["alpha"])
    t = lambda a: None if a is None else torch.from_numpy(a)
    sim_t = fs.conv2d(t(x), t(w), t(bias), t(post), gm["stride"], gm["pad"], gm["dil"], gm["alpha"], gm["center"]).numpy()
    scale = max(1e-6, float(np.abs(sim_t).max()))
    assert got.shape == sim_t.shape == sim_c.shape
    assert np.abs(got - sim_t).max() <= 1e-5 * scale
    assert np.abs(sim_c - sim_t).max() <= 1e-5 * scale


@settings(max_examples=40, deadline=None, derandomize=True)
@given(st.integers(1, 3), st.integers(1, 200), st.integers(1, 6), st.integers(1, 6), st.integers(0, 2 ** 20))
def test_pack_act_bits_are_the_sign_and_nonzero_planes(n, c, h, w, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, c, h, w)).astype(np.float32)
    x[rng.random(x.shape) < 0.4] = 0.0
    x.reshape(-1)[::7] = -0.0
    bits = co.pack_act(x)                                   # [n, chunks, h, w, {s_lo, s_hi, m_lo, m_hi}]
    assert bits.shape == (n, (c + 63) // 64, h, w, 4)
    for ch in range(c):
        word, bit = (ch % 64) // 32, ch % 32
        s = (bits[:, ch // 64, :, :, word] >> np.uint32(bit)) & 1
        m = (bits[:, ch // 64, :, :, 2 + word] >> np.uint32(bit)) & 1
        assert np.array_equal(s.astype(bool), x[:, ch] > 0)
        assert np.array_equal(m.astype(bool), x[:, ch] != 0)
    r = c % 64                                              # channels beyond c are "zero": no sign bit, no mask bit
    if r:
        last = bits[:, -1]
        if r <= 32:
            assert not np.any(last[..., 1]) and not np.any(last[..., 3])
            if r < 32:
                assert not np.any(last[..., 0] >> np.uint32(r)) and not np.any(last[..., 2] >> np.uint32(r))
        else:
            assert not np.any(last[..., 1] >> np.uint32(r - 32)) and not np.any(last[..., 3] >> np.uint32(r - 32))


@st.composite
def fused_configs(draw):
    res = draw(st.sampled_from([None, "pre", "post"]))
    nx = draw(st.booleans())
    act = draw(st.sampled_from([0, 1, 2]))
    # the Hierarchical-Block form: planes of relu(nx(.)) taken before a post-activation shortcut add
    hblock = res == "post" and nx and act == 0 and draw(st.booleans())
    return dict(cin=draw(st.sampled_from([3, 64, 70, 128])), cout=draw(st.sampled_from([8, 32, 64, 96])),
                h=draw(st.integers(3, 7)), w=draw(st.integers(3, 9)), k=draw(st.sampled_from([1, 3])),
                stride=draw(st.integers(1, 2)), bn=draw(st.booleans()), bias=draw(st.booleans()), post=draw(st.booleans()),
                res=res, act=act, nx=nx, hblock=hblock, seed=draw(st.integers(0, 2 ** 20)))


@settings(max_examples=60, deadline=None, derandomize=True)
@given(fused_configs())
def test_fused_epilogue_oracle_equals_the_module_sequence(fc):
    """The oracle's fused epilogue (struct bnn_epilogue) against the torch module sequence of the reference's blocks --
    conv (float simulation) -> BatchNorm(eval, folded) -> (+shortcut) -> ReLU / PReLU -> (+shortcut) -- on random
    configurations; the emitted planes are the bit-pack of the oracle's own fp32 result (knife-edge ulps aside)."""
    rng = np.random.default_rng(fc["seed"])
    k, pad = fc["k"], fc["k"] // 2
    x = np.maximum(rng.standard_normal((2, fc["cin"], fc["h"], fc["w"])), 0).astype(np.float32)
    w = (rng.standard_normal((fc["cout"], fc["cin"], k, k)) * 0.05).astype(np.float32)
    g = co.geom(2, fc["cin"], fc["h"], fc["w"], fc["cout"], k, k, (fc["stride"],) * 2, (pad, pad), (1, 1))
    ho, wo = co.out_hw(g)
    c = fc["cout"]
    bias = (rng.standard_normal(c) * 0.3).astype(np.float32) if fc["bias"] else None
    post = (0.5 + rng.random(c)).astype(np.float32) if fc["post"] else None
    bn = ((0.5 + rng.random(c)).astype(np.float32) * 3, (rng.standard_normal(c) * 0.3).astype(np.float32)) if fc["bn"] else None
    residual = rng.standard_normal((2, c, ho, wo)).astype(np.float32) if fc["res"] else None
    slope = (rng.random(c) * 0.5).astype(np.float32) if fc["act"] == 2 else None
    nx = ((0.5 + rng.random(c)).astype(np.float32), (rng.standard_normal(c) * 0.2).astype(np.float32)) if fc["nx"] else None
    wb, alpha, nz = co.pack_weight(w, True, True)
    if nz:
        return
    out, bits = co.bconv2d_fused(co.pack_act(x), wb, g, scale=alpha, bias=bias, post=post, bn=bn, residual=residual,
                                 residual_after_act=fc["res"] == "post", act=fc["act"], act_slope=slope, want_out=True,
                                 want_bits=True, nx=nx, nx_relu=fc["hblock"], bits_before_residual=fc["hblock"])
    t = lambda a: None if a is None else torch.from_numpy(a)
    v = lambda a: t(a).view(1, -1, 1, 1)
    y = fs.conv2d(t(x), t(w), t(bias), t(post), (fc["stride"],) * 2, (pad, pad), (1, 1), True, True)
    if bn is not None:
        y = y * v(bn[0]) + v(bn[1])
    if fc["res"] == "pre":
        y = y + t(residual)
    if fc["act"] == 1:
        y = torch.relu(y)
    elif fc["act"] == 2:
        y = torch.nn.functional.prelu(y, t(slope))
    y_before = y
    if fc["res"] == "post":
        y = y + t(residual)
    scale = max(1e-6, float(y.abs().max()))
    assert np.abs(out - y.numpy()).max() <= 1e-5 * scale
    if fc["hblock"]:
        nb = torch.relu(y_before * v(nx[0]) + v(nx[1])).numpy()
        assert (bits != co.pack_act(nb)).sum() <= 2
        return
    want_bits = co.pack_act(out, pre_scale=None if nx is None else nx[0], pre_shift=None if nx is None else nx[1])
    assert (bits != want_bits).sum() <= (1 if nx is not None else 0)


@st.composite
def conv_geoms(draw):
    k = draw(st.sampled_from([1, 3, 5, 7]))
    return dict(n=draw(st.integers(1, 300)), c_in=draw(st.integers(1, 2048)), c_out=draw(st.integers(1, 2048)),
                h=draw(st.integers(k, 120)), w=draw(st.integers(k, 300)), k=k, stride=draw(st.integers(1, 2)),
                pad=draw(st.integers(0, k // 2)), flags=draw(st.sampled_from([0, 2])))


@settings(max_examples=150, deadline=None, derandomize=True)
@given(conv_geoms())
def test_tile_planner_invariants(gm):
    """bnn_conv_plan (host only): whatever geometry comes in, the chosen plan covers every output pixel and channel,
    fits shared memory, keeps whole pixel groups per tile row and uses an instance that exists."""
    from bnn_b200 import native
    g = native.ConvGeom(gm["n"], gm["c_in"], gm["h"], gm["w"], gm["c_out"], gm["k"], gm["k"], gm["stride"], gm["stride"],
                        gm["pad"], gm["pad"], 1, 1)
    ho = (gm["h"] + 2 * gm["pad"] - gm["k"]) // gm["stride"] + 1
    wo = (gm["w"] + 2 * gm["pad"] - gm["k"]) // gm["stride"] + 1
    try:
        pl = native.conv_plan(g, gm["flags"], 148)
    except native.NativeError:
        # refusal is legitimate only if even the smallest tile (4 pixels x 32 channels, one row) exceeds shared memory:
        # plan_smem() of bconv.cu for P = 4, C = 1, TH = 1, TW = 4, 7 warps
        nch, k = (gm["c_in"] + 63) // 64, gm["k"]
        act = nch * k * ((4 - 1) * gm["stride"] + k) * 16
        smallest = 128 + ((act + 127) & ~127) + nch * k * k * 256 + 7 * 32 * 5 * 4 + 5 * 32 * 4 + 4 * 4
        assert smallest > 220 * 1024 or nch * k * k * 256 + 16 * 1024 > 220 * 1024
        return
    assert pl["P"] in (8, 7, 4) and pl["C"] in (4, 2, 1) and pl["warps"] in (2, 4, 7, 8)
    assert pl["TW"] % pl["P"] == 0 and pl["groups"] == pl["TH"] * (pl["TW"] // pl["P"])
    assert pl["smem"] <= 220 * 1024
    assert pl["kw_inst"] == (gm["k"] if gm["k"] in (1, 3) and not (gm["k"] == 1 and gm["stride"] == 2) else 0)
    assert pl["csa"] == int(gm["flags"] == 0 and pl["kw_inst"] in (1, 3))
    tiles_per_image = pl["units"] // gm["n"]
    assert pl["units"] == tiles_per_image * gm["n"]
    assert tiles_per_image * pl["TH"] * pl["TW"] >= ho * wo                       # every output pixel is in some tile
    assert pl["channel_tiles"] * 32 * pl["C"] >= gm["c_out"]                        # every output channel in some block
    assert (pl["channel_tiles"] - 1) * 32 * pl["C"] < gm["c_out"]                   # ... and no empty channel tile


def test_small_cta_plans_are_listed_only_when_all_of_them_fit():
    """Host only (bnn_conv_plan_list): 4- and 2-warp CTAs are candidates with proportionally more CTAs resident -- the same
    warps per SM as the 8-warp form of the instance -- and only when shared memory holds all of them."""
    from bnn_b200 import native

    def plans(n, ci, h, w, co, k, s, p):
        return native.conv_plan_list(native.ConvGeom(n, ci, h, w, co, k, k, s, s, p, p, 1, 1), 0)

    for geom in ((256, 64, 56, 56, 64, 3, 1, 1), (256, 512, 7, 7, 512, 3, 1, 1), (128, 256, 14, 14, 1024, 1, 1, 0)):
        for p in plans(*geom):
            assert p["warps"] in (8, 7, 4, 2)
            if p["warps"] < 7:
                one_by_one = p["kw_inst"] == 1 and p["P"] * p["C"] <= 16
                regs = 80 if one_by_one else (96 if p["C"] <= 2 else 128)       # bconv_regs() of bconv_kernel.cuh
                warps_per_sm = 24 if one_by_one else 2048 // regs
                resident = min(32, warps_per_sm // p["warps"])
                assert resident * p["smem"] <= 226 * 1024, p
    # layer1 (4.6 KB of weights per CTA): every (P, C) family has 4- and 2-warp forms; layer4 with C = 4 (73 KB): none
    l1 = {(p["P"], p["C"], p["warps"]) for p in plans(256, 64, 56, 56, 64, 3, 1, 1)}
    assert {(8, 2, 4), (8, 2, 2), (7, 2, 4)} <= l1
    l4 = {(p["P"], p["C"], p["warps"]) for p in plans(256, 512, 7, 7, 512, 3, 1, 1)}
    assert not any(c == 4 and w < 7 for (_, c, w) in l4) and (7, 2, 4) in l4
