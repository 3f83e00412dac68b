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


@settings(max_examples=60, deadline=None)
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


@settings(max_examples=40, deadline=None)
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
