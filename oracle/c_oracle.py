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


def pack_act(x):
    """x: float32 [n,c,h,w] (any numpy strides) -> (abits uint32 [n,chunks,h,w,4], cnt uint32 [n,h,w])."""
    x = np.asarray(x, dtype=np.float32)
    n, c, h, w = x.shape
    sn, sc, sh, sw = (s // 4 for s in x.strides)
    nch = (c + 63) // 64
    abits = np.zeros((n, nch, h, w, 4), np.uint32)
    cnt = np.zeros((n, h, w), np.uint32)
    lib().orc_pack_act(_p(x), c_int64(sn), c_int64(sc), c_int64(sh), c_int64(sw), n, c, h, w, _p(abits), _p(cnt))
    return abits, cnt


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


def bconv2d(abits, cnt, wbits, scale, bias, post, g: Geom, out_strides=None):
    ho, wo = out_hw(g)
    out = np.zeros((g.n, g.c_out, ho, wo), np.float32)
    if out_strides is None:
        out_strides = tuple(s // 4 for s in out.strides)
    scale, bias, post = _f32(scale), _f32(bias), _f32(post)
    lib().orc_bconv2d(_p(abits), _p(cnt), _p(wbits), _p(scale), _p(bias), _p(post), ctypes.byref(g), _p(out),
                      *(c_int64(s) for s in out_strides))
    return out


def bconv2d_dot(abits, cnt, wbits, g: Geom):
    ho, wo = out_hw(g)
    dot = np.zeros((g.n, g.c_out, ho, wo), np.int32)
    lib().orc_bconv2d_dot(_p(abits), _p(cnt), _p(wbits), ctypes.byref(g), _p(dot))
    return dot
