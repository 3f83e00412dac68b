"""Forced-plan sweep: EVERY compiled instance of bconv_kernel<P, C, KW, SW, MODE, EPI> against the oracle, bit for bit.

Which tile plan a launch uses is normally the autotuner's choice, so the instances a timed run picks
(e.g. <7,2,3,1,1,3> at bs 256) were not pinned by any test.  ``bnn_bconv2d_fused_fwd_plan`` forces the plan;
this file walks (P, C) in {8,7,4} x {4,2,1} for each unrolled kernel-width / stride instance, both inner-loop
modes (carry-save on / off) and the five epilogue instances, and ``test_sweep_covers_every_instance`` checks on the
host that the sweep really reaches every instance the library compiles.
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle as co

from bnn_b200 import functional as BF
from bnn_b200 import native

DEV = "cuda:0"

GEOMS = [
    dict(name="k3s1", cin=128, cout=128, hw=(5, 56), k=3, stride=1, pad=1),          # instance <3,1>
    dict(name="k3s2", cin=64, cout=128, hw=(9, 112), k=3, stride=2, pad=1),          # instance <3,2>
    dict(name="k1s1", cin=256, cout=128, hw=(4, 56), k=1, stride=1, pad=0),          # instance <1,1> (chunk-triple CSA + 1)
    dict(name="k5", cin=64, cout=128, hw=(6, 56), k=5, stride=1, pad=2),             # generic instance <0,0>
    dict(name="k1s2", cin=128, cout=128, hw=(6, 111), k=1, stride=2, pad=0),         # generic instance, strided 1x1
]
# name -> (epilogue fields, flags).  Which EPI instance each one selects is asserted below.
EPIS = {
    "ref": (dict(bias=True, post=True), 0),                                                              # EPI 0, CSA
    "ref_nocsa": (dict(bias=True, post=True), native.F_NO_CSA),                                          # EPI 0, no CSA
    "fused_nchw": (dict(bn=True, act=2, res="pre", bits=True, out=True, nx=True, cl=False), 0),          # EPI 1
    "fused_cl_hblock": (dict(res="post", bits=True, out=True, nx=True, nx_relu=True, bits_pre=True), 0),  # EPI 2
    "lean_relu": (dict(bn=True, act=1, res="pre", bits=True, out=True), 0),                              # EPI 3 (or 2)
    "lean_prelu": (dict(act=2, res="post", bits=True, out=True, nx=True, bias=True), 0),                 # EPI 4 (or 2)
}
EXPECT_EPI = {"ref": 0, "ref_nocsa": 0, "fused_nchw": 1, "fused_cl_hblock": 2, "lean_relu": 3, "lean_prelu": 4}


def _geom(gm, n=2):
    h, w = gm["hw"]
    return native.ConvGeom(n, gm["cin"], h, w, gm["cout"], gm["k"], gm["k"], gm["stride"], gm["stride"], gm["pad"],
                           gm["pad"], 1, 1)


def _families(gm, flags):
    fams = []
    for p in native.conv_plan_list(_geom(gm), flags):
        key = (p["P"], p["C"])
        if key not in fams:
            fams.append(key)
    return fams


def _dummy_epilogue(spec, cl=True):
    """struct bnn_epilogue with non-NULL placeholders: only bnn_conv_instance (host side, no dereference) sees it."""
    ep = native.Epilogue()
    one = 16
    ep.scale = one
    if spec.get("bn"):
        ep.bn_scale = ep.bn_shift = one
    if spec.get("res"):
        ep.residual, ep.rstride_c = one, (1 if spec.get("cl", True) else 7)
        ep.residual_after_act = int(spec["res"] == "post")
    ep.act = spec.get("act", 0)
    if ep.act == 2:
        ep.act_slope = one
    if spec.get("out", True):
        ep.out, ep.ostride_c = one, (1 if spec.get("cl", True) else 7)
    if spec.get("bits"):
        ep.out_bits = one
    if spec.get("nx"):
        ep.nx_scale = ep.nx_shift = one
    ep.nx_relu, ep.bits_before_residual = int(bool(spec.get("nx_relu"))), int(bool(spec.get("bits_pre")))
    return ep


def _instances(gm, ename):
    spec, flags = EPIS[ename]
    fused = ename not in ("ref", "ref_nocsa")
    out = {}
    for (P, C) in _families(gm, flags):
        ep = _dummy_epilogue(spec) if fused else _dummy_epilogue(dict(out=True, cl=False))
        try:
            inst = native.conv_instance(_geom(gm), ep, flags, P, C)
        except native.NativeError:
            continue                       # no instance for this family (lean epilogues need C >= 2 -> falls to EPI 2 first)
        out[(P, C)] = inst
    return out


def test_sweep_covers_every_instance():
    """Host only: the (geometry x epilogue x family) sweep below reaches every instance bconv_inst.cu compiles."""
    seen = set()
    for gm in GEOMS:
        for ename in EPIS:
            for inst in _instances(gm, ename).values():
                seen.add((inst["P"], inst["C"], inst["kw_inst"], inst["stride_inst"], inst["csa"], inst["epi"]))
    want = set()
    for epi in range(5):
        for (kw, sw, modes) in ((3, 1, (0, 1)), (3, 2, (0, 1)), (1, 1, (0, 1)), (0, 0, (0,))):
            for mode in modes:
                if mode == 0 and epi != 0 and kw in (3,):
                    continue                               # fused epilogues always run the carry-save loop for 3-wide rows
                if mode == 0 and epi != 0 and kw == 1:
                    continue
                for P in (8, 7, 4):
                    for C in ((4, 2, 1) if epi < 3 else (4, 2)):
                        want.add((P, C, kw, sw, mode, epi))
    missing = sorted(want - seen)
    assert not missing, f"instances never exercised by the sweep: {missing[:10]} ... ({len(missing)})"


def _inputs(gm, ename):
    spec, _ = EPIS[ename]
    rng = np.random.default_rng(777 + 31 * [g["name"] for g in GEOMS].index(gm["name"]) + list(EPIS).index(ename))
    n, (h, w), k = 2, gm["hw"], gm["k"]
    x = np.maximum(rng.standard_normal((n, gm["cin"], h, w)), 0).astype(np.float32)
    x.reshape(-1)[::19] *= -1.0                                  # some negatives: s != m
    wt = (rng.standard_normal((gm["cout"], gm["cin"], k, k)) * 0.05).astype(np.float32)
    g = co.geom(n, gm["cin"], h, w, gm["cout"], k, k, (gm["stride"],) * 2, (gm["pad"],) * 2, (1, 1))
    ho, wo = co.out_hw(g)
    c = gm["cout"]
    f32 = lambda a: a.astype(np.float32)
    return dict(
        x=x, w=wt, g=g,
        bias=f32(rng.standard_normal(c) * 0.3) if spec.get("bias") else None,
        post=f32(0.5 + rng.random(c)) if spec.get("post") else None,
        bn=(f32((0.5 + rng.random(c)) * 3), f32(rng.standard_normal(c) * 0.3)) if spec.get("bn") else None,
        residual=f32(rng.standard_normal((n, c, ho, wo))) if spec.get("res") else None,
        res_after=spec.get("res") == "post", act=spec.get("act", 0),
        slope=f32(rng.random(c) * 0.5) if spec.get("act", 0) == 2 else None,
        nx=(f32(0.5 + rng.random(c)), f32(rng.standard_normal(c) * 0.2)) if spec.get("nx") else None,
        nx_relu=bool(spec.get("nx_relu")), bits_pre=bool(spec.get("bits_pre")))


def _d(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.gpu
@pytest.mark.parametrize("ename", list(EPIS))
@pytest.mark.parametrize("gm", GEOMS, ids=[g["name"] for g in GEOMS])
def test_every_tile_family_bit_exact_vs_oracle(gm, ename):
    spec, flags = EPIS[ename]
    d = _inputs(gm, ename)
    g = d["g"]
    act = BF.pack_activations(_d(d["x"]))
    wts = BF.pack_weights(_d(d["w"]), True, True)
    alpha = wts.alpha.cpu().numpy()                       # device alpha on both sides (differs in the last bit at most)
    ab = co.pack_act(d["x"])
    wb, _, nz = co.pack_weight(d["w"], True, True)
    assert nz == 0
    fused = ename not in ("ref", "ref_nocsa")
    cl = spec.get("cl", True)
    if fused:
        want_out, want_bits = co.bconv2d_fused(ab, wb, g, scale=alpha, bias=d["bias"], post=d["post"], bn=d["bn"],
                                               residual=d["residual"], residual_after_act=d["res_after"], act=d["act"],
                                               act_slope=d["slope"], want_out=True, want_bits=True, nx=d["nx"],
                                               nx_relu=d["nx_relu"], bits_before_residual=d["bits_pre"])
    else:
        want_out, want_bits = co.bconv2d(ab, wb, alpha, d["bias"], d["post"], g), None
    insts = _instances(gm, ename)
    assert insts
    pair = lambda p: None if p is None else (_d(p[0]), _d(p[1]))
    res = _d(d["residual"])
    if res is not None and cl:
        res = res.contiguous(memory_format=torch.channels_last)
    stride, pad = (g.stride_h, g.stride_w), (g.pad_h, g.pad_w)
    ran = 0
    for (P, C), inst in insts.items():
        if inst["epi"] != EXPECT_EPI[ename]:
            # e.g. lean epilogues with C == 1 do not exist; those families run EPI 2 -- still compared below
            assert ename.startswith("lean") and inst["epi"] == 2, (ename, P, C, inst)
        for warps in (0, 7, 4):                     # best / 7-warp / 4-warp (half-size CTA) candidate of the family
            try:
                out, bits = BF.bconv2d_fused(act, wts, bias=_d(d["bias"]), post=_d(d["post"]), bn=pair(d["bn"]),
                                             residual=res, residual_after_act=d["res_after"], activation=d["act"],
                                             act_slope=_d(d["slope"]), want_out=True, want_bits=fused and spec.get("bits", False),
                                             nx=pair(d["nx"]), stride=stride, padding=pad, flags=flags, channels_last=fused and cl,
                                             nx_relu=d["nx_relu"], bits_before_residual=d["bits_pre"], plan=(P, C, 0, warps))
            except native.NativeError:
                assert warps != 0                          # a 7- or 4-warp variant of this family may not exist
                continue
            ran += 1
            assert np.array_equal(out.cpu().numpy(), want_out), (gm["name"], ename, P, C, warps)
            if bits is not None:
                assert np.array_equal(bits.bits.cpu().numpy().view(np.uint32), want_bits), (gm["name"], ename, P, C, warps)
    assert ran >= len(insts)
