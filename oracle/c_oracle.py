"""numpy front-end of oracle/bnn_oracle.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes
from ctypes import POINTER, c_float, c_int, c_int32, c_int64, c_uint32, c_void_p

import numpy as np

from . import build as _build


class Geom(ctypes.Structure):
    _fields_ = [(k, c_int32) for k in (
        "n", "c_in", "h", "w", "c_out", "kh", "kw", "stride_h", "stride_w", "pad_h", "pad_w", "dil_h", "dil_w")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.orc_act_units.restype = c_int64
        _lib.orc_weight_words.restype = c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def out_hw(g: Geom):
    return lib().orc_out_h(ctypes.byref(g)), lib().orc_out_w(ctypes.byref(g))


def geom(n, c_in, h, w, c_out, kh, kw, stride=(1, 1), pad=(0, 0), dil=(1, 1)) -> Geom:
    return Geom(n, c_in, h, w, c_out, kh, kw, stride[0], stride[1], pad[0], pad[1], dil[0], dil[1])


def floatsim_conv2d(x, w, bias, post, g: Geom, center: bool, compute_alpha: bool):
    x, w, bias, post = _f32(x), _f32(w), _f32(bias), _f32(post)
    ho, wo = out_hw(g)
    out = np.empty((g.n, g.c_out, ho, wo), np.float32)
    rc = lib().orc_floatsim_conv2d(_p(x), _p(w), _p(bias), _p(post), ctypes.byref(g), int(center),
                                   int(compute_alpha), _p(out))
    assert rc == 0
    return out


def floatsim_linear(x, w, bias, post, center: bool, compute_alpha: bool):
    x, w, bias, post = _f32(x), _f32(w), _f32(bias), _f32(post)
    rows, in_f = x.shape
    out = np.empty((rows, w.shape[0]), np.float32)
    rc = lib().orc_floatsim_linear(_p(x), _p(w), _p(bias), _p(post), rows, in_f, w.shape[0], int(center),
                                   int(compute_alpha), _p(out))
    assert rc == 0
    return out


def pack_act(x, pool=0, ceil_mode=True, pre_scale=None, pre_shift=None, pre_relu=False):
    """x: float32 [n,c,h,w] (any numpy strides) -> abits uint32 [n,chunks,ho,wo,4]."""
    x = np.asarray(x, dtype=np.float32)
    n, c, h, w = x.shape
    sn, sc, sh, sw = (s // 4 for s in x.strides)
    nch = (c + 63) // 64
    k = pool if pool > 1 else 1
    ho = (-(-h // k) if ceil_mode else h // k) if pool > 1 else h
    wo = (-(-w // k) if ceil_mode else w // k) if pool > 1 else w
    abits = np.zeros((n, nch, ho, wo, 4), np.uint32)
    pre_scale, pre_shift = _f32(pre_scale), _f32(pre_shift)
    lib().orc_pack_act(_p(x), c_int64(sn), c_int64(sc), c_int64(sh), c_int64(sw), n, c, h, w, int(pool),
                       int(ceil_mode), _p(pre_scale), _p(pre_shift), int(pre_relu), _p(abits))
    return abits


def pack_weight(w, center: bool, compute_alpha: bool):
    """w: [c_out,c_in,kh,kw] -> (wbits uint32 [c_out/32, ksteps, 32, 2], alpha float32 [c_out], n_zero)."""
    w = _f32(w)
    if w.ndim == 2:
        w = w[:, :, None, None]
    elif w.ndim == 3:
        w = w[:, :, None, :]
    w = np.ascontiguousarray(w)
    c_out, c_in, kh, kw = w.shape
    nk = ((c_in + 63) // 64) * kh * kw
    wbits = np.zeros(((c_out + 31) // 32, nk, 32, 2), np.uint32)
    alpha = np.zeros((c_out,), np.float32)
    nz = c_int32(0)
    rc = lib().orc_pack_weight(_p(w), c_out, c_in, kh, kw, int(center), int(compute_alpha), _p(wbits), _p(alpha),
                               ctypes.byref(nz))
    assert rc == 0
    return wbits, alpha, int(nz.value)


def bconv2d(abits, wbits, scale, bias, post, g: Geom, out_strides=None):
    ho, wo = out_hw(g)
    out = np.zeros((g.n, g.c_out, ho, wo), np.float32)
    if out_strides is None:
        out_strides = tuple(s // 4 for s in out.strides)
    scale, bias, post = _f32(scale), _f32(bias), _f32(post)
    lib().orc_bconv2d(_p(abits), _p(wbits), _p(scale), _p(bias), _p(post), ctypes.byref(g), _p(out),
                      *(c_int64(s) for s in out_strides))
    return out


def bconv2d_dot(abits, wbits, g: Geom):
    ho, wo = out_hw(g)
    dot = np.zeros((g.n, g.c_out, ho, wo), np.int32)
    lib().orc_bconv2d_dot(_p(abits), _p(wbits), ctypes.byref(g), _p(dot))
    return dot


class Epilogue(ctypes.Structure):
    _fields_ = [("scale", c_void_p), ("bias", c_void_p), ("post", c_void_p), ("bn_scale", c_void_p),
                ("bn_shift", c_void_p), ("residual", c_void_p), ("rn", c_int64), ("rc", c_int64), ("rh", c_int64),
                ("rw", c_int64), ("residual_after_act", c_int32), ("act", c_int32), ("act_slope", c_void_p),
                ("out", c_void_p), ("on", c_int64), ("oc", c_int64), ("oh", c_int64), ("ow", c_int64),
                ("out_bits", c_void_p), ("nx_scale", c_void_p), ("nx_shift", c_void_p), ("nx_relu", c_int32),
                ("bits_before_residual", c_int32)]


def bconv2d_fused(abits, wbits, g: Geom, scale=None, bias=None, post=None, bn=None, residual=None,
                  residual_after_act=False, act=0, act_slope=None, want_out=True, want_bits=False, nx=None,
                  nx_relu=False, bits_before_residual=False):
    """Fused epilogue (struct bnn_epilogue). bn / nx: (scale, shift) pairs. Returns (out or None, bits or None)."""
    ho, wo = out_hw(g)
    keep = [_f32(a) for a in (scale, bias, post, None if bn is None else bn[0], None if bn is None else bn[1],
                              residual, act_slope, None if nx is None else nx[0], None if nx is None else nx[1])]
    scale, bias, post, bns, bnh, residual, act_slope, nxs, nxh = keep
    out = np.zeros((g.n, g.c_out, ho, wo), np.float32) if want_out else None
    bits = np.zeros((g.n, (g.c_out + 63) // 64, ho, wo, 4), np.uint32) if want_bits else None
    e = Epilogue()
    e.scale, e.bias, e.post, e.bn_scale, e.bn_shift = (_v(a) for a in (scale, bias, post, bns, bnh))
    if residual is not None:
        e.residual = _v(residual)
        e.rn, e.rc, e.rh, e.rw = (s // 4 for s in residual.strides)
    e.residual_after_act, e.act, e.act_slope = int(residual_after_act), int(act), _v(act_slope)
    if out is not None:
        e.out = _v(out)
        e.on, e.oc, e.oh, e.ow = (s // 4 for s in out.strides)
    e.out_bits, e.nx_scale, e.nx_shift = _v(bits), _v(nxs), _v(nxh)
    e.nx_relu, e.bits_before_residual = int(nx_relu), int(bits_before_residual)
    lib().orc_bconv2d_fused(_p(abits), _p(wbits), ctypes.byref(g), ctypes.byref(e))
    return out, bits


def _v(a):
    return None if a is None else a.ctypes.data


def stem(x, w, bn_scale, bn_shift, nx=None):
    """ResNet stem (conv7x7/2 + folded BN + ReLU + maxpool3/2/1): returns (out NHWC [n,hp,wp,64], bits [n,1,hp,wp,4])."""
    x, w = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(w, np.float32)
    n, _, h, wd = x.shape
    hc, wc = (h + 6 - 7) // 2 + 1, (wd + 6 - 7) // 2 + 1
    hp, wp = (hc + 2 - 3) // 2 + 1, (wc + 2 - 3) // 2 + 1
    out = np.zeros((n, hp, wp, 64), np.float32)
    bits = np.zeros((n, 1, hp, wp, 4), np.uint32)
    g, hh = _f32(bn_scale), _f32(bn_shift)
    nxs, nxh = (None, None) if nx is None else (_f32(nx[0]), _f32(nx[1]))
    lib().orc_stem(_p(x), n, h, wd, _p(w), _p(g), _p(hh), _p(nxs), _p(nxh), _p(out), _p(bits))
    return out, bits
