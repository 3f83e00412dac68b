"""Tensor-level wrappers over the C ABI: allocate outputs with torch, pass raw pointers and the
current CUDA stream.  torch is plumbing here (device memory + streams); all arithmetic of the
binarized path happens in the kernels of csrc/.
"""
import ctypes
import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import native
from .native import ConvGeom


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda_f32(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise native.NativeError(f"{name} must be a CUDA tensor (got {t.device}); the B200 path has no CPU fallback")
    if t.dtype != torch.float32:
        raise native.NativeError(f"{name} must be float32 (got {t.dtype})")


@dataclass
class PackedActivations:
    """Sign/mask planes of one activation tensor (layout: include/bnn_b200.h)."""
    bits: torch.Tensor   # int32 [n, chunks, h, w, 4]
    n: int
    c: int
    h: int
    w: int


@dataclass
class PackedWeights:
    """Sign planes + XNOR alpha of one weight tensor."""
    bits: torch.Tensor       # int32 [c_out/32, ksteps, 32, 2]
    alpha: torch.Tensor      # float32 [c_out]
    n_zero: int              # exactly-zero (centred) weights: sign 0 in the reference (ternary tensor)
    c_out: int
    c_in: int
    kh: int
    kw: int
    hi: Optional["PackedWeights"] = None     # ternary tensors: the same planes with the zeros packed as +1 (bits: as -1)


def pack_activations(x: torch.Tensor, linear_rows: bool = False, pre: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                     pool: int = 0, ceil_mode: bool = True, pre_relu: bool = False) -> PackedActivations:
    """``BasicInputBinarizer`` (reference bnn/ops.py:151-152) as a bit-pack.

    ``x`` is [n,c,h,w] (any strides) or, with ``linear_rows``, [rows, features] which is packed
    as n=1, h=1, w=rows so that ``blinear`` can treat rows as pixels.  ``pre=(scale, shift)`` folds a
    per-channel affine (eval BatchNorm) in front of the sign, ``pre_relu`` a ReLU between the two; ``pool=k``
    first applies AvgPool2d(k, k, ceil_mode, count_include_pad=False)."""
    _require_cuda_f32(x, "input")
    if linear_rows:
        rows, feat = x.shape
        n, c, h, w = 1, feat, 1, rows
        sn, sc, sh, sw = 0, x.stride(1), 0, x.stride(0)
    else:
        n, c, h, w = x.shape
        sn, sc, sh, sw = x.stride()
    nch = (c + 63) // 64
    ho, wo = h, w
    if pool > 1:
        ho = -(-h // pool) if ceil_mode else h // pool
        wo = -(-w // pool) if ceil_mode else w // pool
    ps, ph = (None, None) if pre is None else (pre[0].data_ptr(), pre[1].data_ptr())
    with torch.cuda.device(x.device):
        bits = torch.empty((n, nch, ho, wo, 4), dtype=torch.int32, device=x.device)
        if pool > 1:
            rc = native.lib().bnn_avgpool_pack_f32(x.data_ptr(), sn, sc, sh, sw, n, c, h, w, pool, int(ceil_mode),
                                                   ps, ph, int(pre_relu), bits.data_ptr(), _stream_ptr(x.device))
        else:
            rc = native.lib().bnn_pack_act_f32(x.data_ptr(), sn, sc, sh, sw, n, c, h, w, ps, ph, int(pre_relu),
                                               bits.data_ptr(), _stream_ptr(x.device))
    native.check(rc, "bnn_pack_act_f32")
    return PackedActivations(bits, n, c, ho, wo)


def avgpool2_pack(x: torch.Tensor, pre: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, pre_relu: bool = False,
                  want_pooled: bool = True):
    """``nn.AvgPool2d(2)`` of a dense channels_last fp32 tensor fused with the bit-pack of the next binarized layer
    (``bnn_avgpool2_pack_cl_f32``): one pass over ``x``.  Returns (pooled channels_last tensor or None,
    PackedActivations of ``sign(pooled * pre[0] + pre[1])``)."""
    _require_cuda_f32(x, "input")
    if x.dim() != 4 or not x.is_contiguous(memory_format=torch.channels_last):
        raise native.NativeError("avgpool2_pack expects a dense channels_last [n,c,h,w] tensor")
    n, c, h, w = x.shape
    ho, wo = h // 2, w // 2
    dev = x.device
    ps, ph = (None, None) if pre is None else (pre[0].data_ptr(), pre[1].data_ptr())
    with torch.cuda.device(dev):
        pooled = (torch.empty((n, c, ho, wo), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
                  if want_pooled else None)
        bits = torch.empty((n, (c + 63) // 64, ho, wo, 4), dtype=torch.int32, device=dev)
        rc = native.lib().bnn_avgpool2_pack_cl_f32(x.data_ptr(), n, c, h, w, ps, ph, int(pre_relu),
                                                   None if pooled is None else pooled.data_ptr(), bits.data_ptr(),
                                                   _stream_ptr(dev))
    native.check(rc, "bnn_avgpool2_pack_cl_f32")
    return pooled, PackedActivations(bits, n, c, ho, wo)


def pack_weights(weight: torch.Tensor, center_weights: bool, compute_alpha: bool) -> PackedWeights:
    """``XNORWeightBinarizer`` (reference bnn/ops.py:129-140) as a prepare-time pack.  A tensor with exactly-zero
    (centred) weights -- sign 0 in the reference, bnn/ops.py:66,136 -- comes back with ``hi`` set: a second set of
    planes with the zeros packed as +1 (``bits`` has them as -1); the ternary dot is the mean of the two binary dots."""
    _require_cuda_f32(weight, "weight")
    w = weight.detach().contiguous()
    if w.dim() == 2:
        c_out, c_in, kh, kw = w.shape[0], w.shape[1], 1, 1
    elif w.dim() == 3:
        c_out, c_in, kh, kw = w.shape[0], w.shape[1], 1, w.shape[2]
    elif w.dim() == 4:
        c_out, c_in, kh, kw = w.shape
    else:
        raise ValueError(f"Expected ndims equal with 2 or 4, but found {w.dim()}")
    nk = ((c_in + 63) // 64) * kh * kw
    with torch.cuda.device(w.device):
        bits = torch.empty(((c_out + 31) // 32, nk, 32, 2), dtype=torch.int32, device=w.device)
        alpha = torch.empty((c_out,), dtype=torch.float32, device=w.device)
        nz = torch.zeros((1,), dtype=torch.int32, device=w.device)
        zero = torch.empty_like(bits)
        rc = native.lib().bnn_pack_weight_ternary_f32(w.data_ptr(), c_out, c_in, kh, kw, int(center_weights),
                                                      int(compute_alpha), bits.data_ptr(), zero.data_ptr(),
                                                      alpha.data_ptr(), nz.data_ptr(), _stream_ptr(w.device))
    native.check(rc, "bnn_pack_weight_ternary_f32")
    n_zero = int(nz.item())
    packed = PackedWeights(bits, alpha, n_zero, c_out, c_in, kh, kw)
    if n_zero:
        packed.hi = PackedWeights(bits | zero, alpha, n_zero, c_out, c_in, kh, kw)
    return packed


def _opt_ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_tuned = set()


def _maybe_tune(act, wts, geom, ep, flags, dev) -> None:
    """First (non-captured) launch of a geometry: let the library time its candidate tile plans."""
    from . import runtime
    key = (tuple(getattr(geom, f) for f, _ in ConvGeom._fields_), bool(ep.bn_scale or ep.residual or ep.act or ep.out_bits
                                                                      or ep.nx_scale), flags & native.F_NO_CSA, dev.index)
    if key in _tuned or not runtime.autotune() or torch.cuda.is_current_stream_capturing():
        return
    rc = native.lib().bnn_bconv2d_tune(act.bits.data_ptr(), wts.bits.data_ptr(), ctypes.byref(geom), ctypes.byref(ep),
                                       flags, 8, _stream_ptr(dev))
    native.check(rc, "bnn_bconv2d_tune")
    _tuned.add(key)


def _out_hw(act, wts, stride, padding, dilation):
    ho = (act.h + 2 * padding[0] - dilation[0] * (wts.kh - 1) - 1) // stride[0] + 1
    wo = (act.w + 2 * padding[1] - dilation[1] * (wts.kw - 1) - 1) // stride[1] + 1
    if ho <= 0 or wo <= 0:
        raise native.NativeError(f"empty output ({ho}x{wo})")
    return ho, wo


_split_cache = {}


def conv_split(geom: ConvGeom, flags: int = 0):
    """(chunks per part, parts) of the split-K decomposition of a geometry (``bnn_conv_split``, host only, cached)."""
    key = (geom.c_in, geom.c_out, geom.kh, geom.kw, geom.stride_h, geom.stride_w, geom.dil_h, geom.dil_w, geom.h, geom.w,
           geom.pad_h, geom.pad_w, flags & native.F_NO_CSA)
    if key not in _split_cache:
        _split_cache[key] = native.conv_split(geom, flags)
    return _split_cache[key]


def needs_general_path(wts: PackedWeights, stride=(1, 1), dilation=(1, 1)) -> bool:
    """True when a layer cannot run as ONE launch of the conv kernel: ternary weights (exact zeros) or a reduction too
    large for a CTA's shared memory (split-K).  Such layers run ``_bconv2d_general`` and are never fused."""
    if wts.hi is not None:
        return True
    probe = ConvGeom(1, wts.c_in, max(8, wts.kh * dilation[0]), max(8, wts.kw * dilation[1]), wts.c_out, wts.kh, wts.kw,
                     stride[0], stride[1], 0, 0, dilation[0], dilation[1])
    try:
        return conv_split(probe)[1] > 1
    except native.NativeError:
        return True


def _bconv2d_general(act: PackedActivations, wts: PackedWeights, geom: ConvGeom, ho: int, wo: int, bias, post, use_alpha,
                     flags, out: torch.Tensor) -> torch.Tensor:
    """Ternary weights and / or split-K: integer dots of every (weight set, chunk range) by ``bnn_bconv2d_partial_fwd``,
    summed (and halved for a ternary pair) and put through the reference epilogue by ``bnn_dot_finish_f32``."""
    dev = act.bits.device
    cpp, nparts = conv_split(geom, flags)
    nch = (geom.c_in + 63) // 64
    sets = [wts] if wts.hi is None else [wts, wts.hi]
    count = act.n * wts.c_out * ho * wo
    lib = native.lib()
    with torch.cuda.device(dev):
        parts = torch.empty((len(sets) * nparts, count), dtype=torch.float32, device=dev)
        i = 0
        for ws in sets:
            for p in range(nparts):
                c0 = p * cpp
                rc = lib.bnn_bconv2d_partial_fwd(act.bits.data_ptr(), ws.bits.data_ptr(), ctypes.byref(geom), c0,
                                                 min(cpp, nch - c0), parts[i].data_ptr(), flags, _stream_ptr(dev))
                native.check(rc, "bnn_bconv2d_partial_fwd")
                i += 1
        on, oc, oh, ow = out.stride()
        rc = lib.bnn_dot_finish_f32(parts.data_ptr(), len(sets) * nparts, len(sets),
                                    wts.alpha.data_ptr() if use_alpha else None, _opt_ptr(bias), _opt_ptr(post),
                                    out.data_ptr(), on, oc, oh, ow, act.n, wts.c_out, ho, wo, _stream_ptr(dev))
    native.check(rc, "bnn_dot_finish_f32")
    return out


def bconv2d(act: PackedActivations, wts: PackedWeights, bias: Optional[torch.Tensor] = None,
            post: Optional[torch.Tensor] = None, stride: Tuple[int, int] = (1, 1),
            padding: Tuple[int, int] = (0, 0), dilation: Tuple[int, int] = (1, 1),
            use_alpha: bool = True, flags: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Packed binary convolution + fused ``(alpha*dot + bias) * post`` epilogue -> fp32 NCHW."""
    if act.c != wts.c_in:
        raise native.NativeError(f"channel mismatch: activations {act.c}, weights {wts.c_in}")
    geom = ConvGeom(act.n, act.c, act.h, act.w, wts.c_out, wts.kh, wts.kw, stride[0], stride[1],
                    padding[0], padding[1], dilation[0], dilation[1])
    ho, wo = _out_hw(act, wts, stride, padding, dilation)
    dev = act.bits.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((act.n, wts.c_out, ho, wo), dtype=torch.float32, device=dev)
        elif tuple(out.shape) != (act.n, wts.c_out, ho, wo) or out.dtype != torch.float32 or not out.is_cuda:
            raise native.NativeError(f"out must be a float32 CUDA tensor of shape {(act.n, wts.c_out, ho, wo)}")
        if wts.hi is not None or conv_split(geom, flags)[1] > 1:
            return _bconv2d_general(act, wts, geom, ho, wo, bias, post, use_alpha, flags, out)
        on, oc, oh, ow = out.stride()
        ep = native.Epilogue()
        ep.scale = wts.alpha.data_ptr() if use_alpha else None
        ep.bias, ep.post, ep.out = _opt_ptr(bias), _opt_ptr(post), out.data_ptr()
        ep.ostride_n, ep.ostride_c, ep.ostride_h, ep.ostride_w = on, oc, oh, ow
        _maybe_tune(act, wts, geom, ep, flags, dev)
        rc = native.lib().bnn_bconv2d_fwd(act.bits.data_ptr(), wts.bits.data_ptr(),
                                          wts.alpha.data_ptr() if use_alpha else None, _opt_ptr(bias),
                                          _opt_ptr(post), out.data_ptr(), on, oc, oh, ow, ctypes.byref(geom),
                                          flags, _stream_ptr(dev))
    native.check(rc, "bnn_bconv2d_fwd")
    return out


def bconv2d_fused(act: PackedActivations, wts: PackedWeights, *, bias=None, post=None, bn=None, residual=None,
                  residual_after_act: bool = False, activation: int = native.ACT_NONE, act_slope=None,
                  want_out: bool = True, want_bits: bool = False, nx=None, stride=(1, 1), padding=(0, 0),
                  dilation=(1, 1), use_alpha: bool = True, flags: int = 0, channels_last: bool = False,
                  out: Optional[torch.Tensor] = None, nx_relu: bool = False, bits_before_residual: bool = False,
                  plan: Optional[Tuple[int, int, int, int]] = None):
    """Binary convolution with the cross-module epilogue of ``struct bnn_epilogue``:
    ``y=(alpha*dot+bias)*post; z=y*bn[0]+bn[1]; (+residual); act; (+residual)`` -> fp32 ``out`` and/or the
    packed planes of ``sign(v*nx[0]+nx[1])`` for the next binarized layer.  Returns (out, PackedActivations).
    ``channels_last`` allocates ``out`` in torch's NHWC memory format: with lanes <-> channels in the kernel a
    warp then stores (and reads the residual) as whole 128-byte lines, no transposition needed.
    ``plan=(P, C, TH, warps)`` forces the tile plan (``bnn_bconv2d_fused_fwd_plan``; TH / warps 0 = best of that
    family) instead of the tuned / modelled one -- the parity suite sweeps every kernel instance with it."""
    if act.c != wts.c_in:
        raise native.NativeError(f"channel mismatch: activations {act.c}, weights {wts.c_in}")
    if wts.hi is not None:
        raise native.NativeError("ternary weights (exact zeros) have no fused lowering; use bconv2d")
    geom = ConvGeom(act.n, act.c, act.h, act.w, wts.c_out, wts.kh, wts.kw, stride[0], stride[1],
                    padding[0], padding[1], dilation[0], dilation[1])
    ho, wo = _out_hw(act, wts, stride, padding, dilation)
    dev = act.bits.device
    ep = native.Epilogue()
    ep.scale = wts.alpha.data_ptr() if use_alpha else None
    ep.bias, ep.post = _opt_ptr(bias), _opt_ptr(post)
    if bn is not None:
        ep.bn_scale, ep.bn_shift = bn[0].data_ptr(), bn[1].data_ptr()
    if residual is not None:
        if tuple(residual.shape) != (act.n, wts.c_out, ho, wo):
            raise native.NativeError(f"residual shape {tuple(residual.shape)} != output {(act.n, wts.c_out, ho, wo)}")
        _require_cuda_f32(residual, "residual")
        ep.residual = residual.data_ptr()
        ep.rstride_n, ep.rstride_c, ep.rstride_h, ep.rstride_w = residual.stride()
    ep.residual_after_act, ep.act, ep.act_slope = int(residual_after_act), int(activation), _opt_ptr(act_slope)
    if nx is not None:
        ep.nx_scale, ep.nx_shift = nx[0].data_ptr(), nx[1].data_ptr()
    ep.nx_relu, ep.bits_before_residual = int(nx_relu), int(bits_before_residual)
    bits = None
    with torch.cuda.device(dev):
        if out is not None:                     # caller-provided view (e.g. a channel slice of a block output)
            if tuple(out.shape) != (act.n, wts.c_out, ho, wo) or out.dtype != torch.float32 or not out.is_cuda:
                raise native.NativeError(f"out must be a float32 CUDA tensor of shape {(act.n, wts.c_out, ho, wo)}")
            want_out = True
        elif want_out:
            out = torch.empty((act.n, wts.c_out, ho, wo), dtype=torch.float32, device=dev,
                              memory_format=torch.channels_last if channels_last else torch.contiguous_format)
        if want_out:
            ep.out = out.data_ptr()
            ep.ostride_n, ep.ostride_c, ep.ostride_h, ep.ostride_w = out.stride()
        if want_bits:
            bits = torch.empty((act.n, (wts.c_out + 63) // 64, ho, wo, 4), dtype=torch.int32, device=dev)
            ep.out_bits = bits.data_ptr()
        if plan is not None:
            rc = native.lib().bnn_bconv2d_fused_fwd_plan(act.bits.data_ptr(), wts.bits.data_ptr(), ctypes.byref(geom),
                                                         ctypes.byref(ep), flags, *[int(v) for v in plan],
                                                         _stream_ptr(dev))
        else:
            _maybe_tune(act, wts, geom, ep, flags, dev)
            rc = native.lib().bnn_bconv2d_fused_fwd(act.bits.data_ptr(), wts.bits.data_ptr(), ctypes.byref(geom),
                                                    ctypes.byref(ep), flags, _stream_ptr(dev))
    native.check(rc, "bnn_bconv2d_fused_fwd")
    packed = None if bits is None else PackedActivations(bits, act.n, wts.c_out, ho, wo)
    return out, packed


def blinear(act: PackedActivations, wts: PackedWeights, bias: Optional[torch.Tensor] = None,
            post: Optional[torch.Tensor] = None, use_alpha: bool = True, flags: int = 0) -> torch.Tensor:
    """Packed binary linear layer: activations packed with ``linear_rows=True`` -> [rows, out]."""
    rows = act.w
    dev = act.bits.device
    geom = ConvGeom(1, act.c, 1, rows, wts.c_out, 1, 1, 1, 1, 0, 0, 1, 1)
    if wts.hi is not None or conv_split(geom, flags)[1] > 1:
        # rows play the role of the image width: out[rows, out] is the [1, out, 1, rows] result with strides (0, 1, 0, out)
        with torch.cuda.device(dev):
            out = torch.empty((rows, wts.c_out), dtype=torch.float32, device=dev)
        view = out.as_strided((1, wts.c_out, 1, rows), (0, 1, 0, wts.c_out))
        _bconv2d_general(act, wts, geom, 1, rows, bias, post, use_alpha, flags, view)
        return out
    with torch.cuda.device(dev):
        out = torch.empty((rows, wts.c_out), dtype=torch.float32, device=dev)
        rc = native.lib().bnn_blinear_fwd(act.bits.data_ptr(), wts.bits.data_ptr(),
                                          wts.alpha.data_ptr() if use_alpha else None, _opt_ptr(bias),
                                          _opt_ptr(post), out.data_ptr(), rows, act.c, wts.c_out, flags,
                                          _stream_ptr(dev))
    native.check(rc, "bnn_blinear_fwd")
    return out


def stem_weight_layout(w: torch.Tensor) -> torch.Tensor:
    """[64,3,7,7] conv weight -> [3,7,7,32,2] with out[ci,kh,kw,l,b] = w[b*32+l,ci,kh,kw] (bnn_stem_fwd's w_t)."""
    return w.detach().permute(1, 2, 3, 0).reshape(3, 7, 7, 2, 32).transpose(3, 4).contiguous()


def stem(x: torch.Tensor, w_t: torch.Tensor, bn: Tuple[torch.Tensor, torch.Tensor], nx=None,
         want_bits: bool = True, flags: int = 0):
    """conv7x7/2 + BatchNorm + ReLU + maxpool3x3/2 of the reference's ResNet stem in one kernel.
    ``x`` [n,3,h,w] contiguous, ``w_t`` the weight repacked to [3,7,7,32,2] (``stem_weight_layout``).  Returns (out, PackedActivations):
    ``out`` is [n,64,hp,wp] in channels_last memory format."""
    _require_cuda_f32(x, "input")
    if x.dim() != 4 or x.shape[1] != 3 or not x.is_contiguous():
        raise native.NativeError(f"stem expects a contiguous [n,3,h,w] tensor, got {tuple(x.shape)}")
    n, _, h, w = x.shape
    hp, wp = ctypes.c_int32(0), ctypes.c_int32(0)
    native.check(native.lib().bnn_stem_out_hw(h, w, ctypes.byref(hp), ctypes.byref(wp)), "bnn_stem_out_hw")
    hp, wp = hp.value, wp.value
    dev = x.device
    with torch.cuda.device(dev):
        out = torch.empty((n, 64, hp, wp), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
        bits = torch.empty((n, 1, hp, wp, 4), dtype=torch.int32, device=dev) if want_bits else None
        rc = native.lib().bnn_stem_fwd(x.data_ptr(), n, h, w, w_t.data_ptr(), bn[0].data_ptr(), bn[1].data_ptr(),
                                       None if nx is None else nx[0].data_ptr(),
                                       None if nx is None else nx[1].data_ptr(), out.data_ptr(),
                                       None if bits is None else bits.data_ptr(), flags, _stream_ptr(dev))
    native.check(rc, "bnn_stem_fwd")
    return out, (None if bits is None else PackedActivations(bits, n, 64, hp, wp))


def shortcut_out_shape(x: torch.Tensor, wts: PackedWeights, pool: int, ceil_mode: bool):
    n, _, h, w = x.shape
    k = max(1, int(pool))
    ho, wo = (-(-h // k), -(-w // k)) if ceil_mode else (h // k, w // k)
    return (n, wts.c_out, ho, wo)


def shortcut(x: torch.Tensor, wts: PackedWeights, pool: int, ceil_mode: bool, *, bias=None, post=None, bn=None,
             use_alpha: bool = True, flags: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """AvgPool(pool) -> sign -> binarized conv1x1 -> BatchNorm of the reference's down-sampling shortcut
    (bnn/models/resnet.py:129-133) in one kernel (bnn_shortcut_fwd).  ``x``: channels_last fp32 [n,c,h,w];
    returns [n,c_out,ho,wo] channels_last.  Bit-identical to ``pack_activations(x, pool=...)`` +
    ``bconv2d_fused(..., bn=bn)``."""
    _require_cuda_f32(x, "input")
    if x.dim() != 4 or x.stride(1) != 1:
        raise native.NativeError("shortcut expects a channels_last [n,c,h,w] tensor (channel stride 1)")
    if wts.kh != 1 or wts.kw != 1 or wts.c_in != x.shape[1]:
        raise native.NativeError(f"shortcut expects a 1x1 conv over {x.shape[1]} channels, got "
                                 f"{wts.kh}x{wts.kw} over {wts.c_in}")
    n, c, h, w = x.shape
    k = max(1, int(pool))
    ho, wo = (-(-h // k), -(-w // k)) if ceil_mode else (h // k, w // k)
    dev = x.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((n, wts.c_out, ho, wo), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
        elif (tuple(out.shape) != (n, wts.c_out, ho, wo) or out.dtype != torch.float32 or not out.is_cuda
              or not out.is_contiguous(memory_format=torch.channels_last)):
            raise native.NativeError(f"out must be a channels_last float32 CUDA tensor of shape {(n, wts.c_out, ho, wo)}")
        rc = native.lib().bnn_shortcut_fwd(x.data_ptr(), x.stride(0), x.stride(2), x.stride(3), n, c, h, w, k,
                                           int(bool(ceil_mode)), wts.bits.data_ptr(), wts.c_out,
                                           wts.alpha.data_ptr() if use_alpha else None, _opt_ptr(bias), _opt_ptr(post),
                                           None if bn is None else bn[0].data_ptr(),
                                           None if bn is None else bn[1].data_ptr(), out.data_ptr(), flags,
                                           _stream_ptr(dev))
    native.check(rc, "bnn_shortcut_fwd")
    return out


STEM_X_LOG2_SCALE = 7          # inputs up to |x| < 511 stay inside the fp16 range (raw 0..255 pixels included)


def stem_mma_weights(w: torch.Tensor):
    """[64,3,7,7] fp32 conv weight -> (fragment buffer, w_log2_scale) for ``stem_mma`` (bnn_stem_mma_pack_weight).
    The scale puts max|w| just below 2^14, well inside the fp16 range.  One host sync (prepare time only)."""
    _require_cuda_f32(w, "stem weight")
    if tuple(w.shape) != (64, 3, 7, 7):
        raise native.NativeError(f"stem weight must be [64,3,7,7], got {tuple(w.shape)}")
    w = w.detach().contiguous()
    wmax = float(w.abs().max())
    if not math.isfinite(wmax):
        raise native.NativeError("stem weight has non-finite values")
    log2_scale = 0 if wmax == 0.0 else max(-60, min(60, 13 - math.frexp(wmax)[1]))
    dev = w.device
    with torch.cuda.device(dev):
        frag = torch.empty(native.lib().bnn_stem_mma_weight_bytes() // 4, dtype=torch.int32, device=dev)
        rc = native.lib().bnn_stem_mma_pack_weight(w.data_ptr(), log2_scale, frag.data_ptr(), _stream_ptr(dev))
    native.check(rc, "bnn_stem_mma_pack_weight")
    return frag, log2_scale


def amax(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """max|x| of a contiguous fp32 tensor as a device scalar (``bnn_amax_f32``; no host synchronisation, capturable):
    the input-range guard of the split-fp16 stem kernels."""
    _require_cuda_f32(x, "input")
    if not x.is_contiguous():
        raise native.NativeError("amax expects a contiguous tensor")
    dev = x.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((1,), dtype=torch.float32, device=dev)
        rc = native.lib().bnn_amax_f32(x.data_ptr(), x.numel(), out.data_ptr(), _stream_ptr(dev))
    native.check(rc, "bnn_amax_f32")
    return out


def _stem_split(kernel: str, x: torch.Tensor, wops, bn, nx, want_bits, x_log2_scale, guard, flags):
    _require_cuda_f32(x, "input")
    if x.dim() != 4 or x.shape[1] != 3 or not x.is_contiguous():
        raise native.NativeError(f"stem expects a contiguous [n,3,h,w] tensor, got {tuple(x.shape)}")
    buf, w_log2_scale = wops
    n, _, h, w = x.shape
    hp, wp = ctypes.c_int32(0), ctypes.c_int32(0)
    native.check(native.lib().bnn_stem_out_hw(h, w, ctypes.byref(hp), ctypes.byref(wp)), "bnn_stem_out_hw")
    hp, wp = hp.value, wp.value
    dev = x.device
    with torch.cuda.device(dev):
        x_amax = None
        if guard is True:
            x_amax = amax(x)
        elif isinstance(guard, torch.Tensor):
            x_amax = guard
        out = torch.empty((n, 64, hp, wp), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
        bits = torch.empty((n, 1, hp, wp, 4), dtype=torch.int32, device=dev) if want_bits else None
        fn = native.lib().bnn_stem_mma_fwd if kernel == "mma" else native.lib().bnn_stem_tc_fwd
        rc = fn(x.data_ptr(), n, h, w, buf.data_ptr(), x_log2_scale, None if x_amax is None else x_amax.data_ptr(),
                w_log2_scale, bn[0].data_ptr(), bn[1].data_ptr(), None if nx is None else nx[0].data_ptr(),
                None if nx is None else nx[1].data_ptr(), out.data_ptr(), None if bits is None else bits.data_ptr(),
                flags, _stream_ptr(dev))
    native.check(rc, f"bnn_stem_{kernel}_fwd")
    return out, (None if bits is None else PackedActivations(bits, n, 64, hp, wp))


def stem_mma(x: torch.Tensor, wfrag, bn: Tuple[torch.Tensor, torch.Tensor], nx=None, want_bits: bool = True,
             x_log2_scale: int = STEM_X_LOG2_SCALE, flags: int = 0, guard=False):
    """``stem`` on the mma.sync path (split-fp16 operands, fp32-level accuracy; see include/bnn_b200.h).
    ``wfrag`` = ``stem_mma_weights(conv.weight)``.  Same outputs as ``stem``.  ``guard=True`` measures max|x| on the
    device first (``amax``) and lets the kernel choose the input scale, so no input magnitude can overflow fp16;
    ``guard=False`` uses ``x_log2_scale`` as given (|x| * 2^x_log2_scale < 65504 is then the caller's promise)."""
    return _stem_split("mma", x, wfrag, bn, nx, want_bits, x_log2_scale, guard, flags)


def stem_tc_weights(w: torch.Tensor):
    """[64,3,7,7] fp32 conv weight -> (B-operand image, w_log2_scale) for ``stem_tc`` (bnn_stem_tc_pack_weight)."""
    _require_cuda_f32(w, "stem weight")
    if tuple(w.shape) != (64, 3, 7, 7):
        raise native.NativeError(f"stem weight must be [64,3,7,7], got {tuple(w.shape)}")
    w = w.detach().contiguous()
    wmax = float(w.abs().max())
    if not math.isfinite(wmax):
        raise native.NativeError("stem weight has non-finite values")
    log2_scale = 0 if wmax == 0.0 else max(-60, min(60, 13 - math.frexp(wmax)[1]))
    dev = w.device
    with torch.cuda.device(dev):
        ops = torch.empty(native.lib().bnn_stem_tc_weight_bytes() // 4, dtype=torch.int32, device=dev)
        rc = native.lib().bnn_stem_tc_pack_weight(w.data_ptr(), log2_scale, ops.data_ptr(), _stream_ptr(dev))
    native.check(rc, "bnn_stem_tc_pack_weight")
    return ops, log2_scale


def u8_log2_scale(mean, istd) -> int:
    """Input scale of the split-fp16 stems for uint8 images normalised as (x - mean) * istd: the largest possible
    |value| is known, so no measuring pass is needed."""
    bound = max(max(abs(0.0 - m), abs(255.0 - m)) * abs(s) for m, s in zip(mean, istd))
    return max(-60, min(60, 15 - math.frexp(float(bound))[1])) if bound > 0 else 0


def stem_tc(x: torch.Tensor, wops, bn: Tuple[torch.Tensor, torch.Tensor], nx=None, want_bits: bool = True,
            x_log2_scale: int = STEM_X_LOG2_SCALE, flags: int = 0, guard=False, pool: bool = True, nx_relu: bool = False,
            nx2=None, nx2_relu: bool = False, u8_norm=None):
    """The fp32 stem on the tcgen05 tensor cores (``bnn_stem_tc_run``, csrc/stem_tc.cu): split-fp16 operands,
    accumulators in tensor memory.  ``wops`` = ``stem_tc_weights(conv.weight)``; ``guard`` as in ``stem_mma``.
    ``pool=False`` drops the max-pool (conv-BN-ReLU, the Hierarchical-Block harness stem); ``nx2`` asks for a second set
    of planes from the same output; ``nx_relu`` / ``nx2_relu`` put a ReLU in front of the respective sign.
    ``u8_norm=(mean[3], istd[3])``: ``x`` is a uint8 [n,h,w,3] image batch, normalised in the kernel as
    ``(x.float() - mean) * istd``.  Returns (out channels_last, planes) or (out, planes, planes2) with ``nx2``."""
    if u8_norm is None:
        _require_cuda_f32(x, "input")
        if x.dim() != 4 or x.shape[1] != 3 or not x.is_contiguous():
            raise native.NativeError(f"stem expects a contiguous [n,3,h,w] tensor, got {tuple(x.shape)}")
        n, _, h, w = x.shape
    else:
        if not x.is_cuda or x.dtype != torch.uint8 or x.dim() != 4 or x.shape[3] != 3 or not x.is_contiguous():
            raise native.NativeError(f"uint8 stem input must be a contiguous CUDA uint8 [n,h,w,3] tensor, got {tuple(x.shape)} {x.dtype}")
        n, h, w, _ = x.shape
    ops, w_log2_scale = wops
    dev = x.device
    lib = native.lib()
    hc, wc = (h + 6 - 7) // 2 + 1, (w + 6 - 7) // 2 + 1
    ho, wo = ((hc + 2 - 3) // 2 + 1, (wc + 2 - 3) // 2 + 1) if pool else (hc, wc)
    p = native.StemTcParams()
    with torch.cuda.device(dev):
        x_amax = None
        if u8_norm is not None:
            mean, istd = [float(v) for v in u8_norm[0]], [float(v) for v in u8_norm[1]]
            p.x_dtype = 1
            p.u8_mean = (ctypes.c_float * 3)(*mean)
            p.u8_istd = (ctypes.c_float * 3)(*istd)
            x_log2_scale = u8_log2_scale(mean, istd)
        elif guard is True:
            x_amax = amax(x)
        elif isinstance(guard, torch.Tensor):
            x_amax = guard
        out = torch.empty((n, 64, ho, wo), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
        bits = torch.empty((n, 1, ho, wo, 4), dtype=torch.int32, device=dev) if want_bits else None
        bits2 = torch.empty((n, 1, ho, wo, 4), dtype=torch.int32, device=dev) if nx2 is not None else None
        p.x, p.n, p.h, p.w, p.w_ops, p.w_log2_scale = x.data_ptr(), n, h, w, ops.data_ptr(), w_log2_scale
        p.x_log2_scale, p.x_amax = x_log2_scale, (None if x_amax is None else x_amax.data_ptr())
        p.bn_scale, p.bn_shift, p.pool = bn[0].data_ptr(), bn[1].data_ptr(), int(pool)
        if nx is not None:
            p.nx_scale, p.nx_shift = nx[0].data_ptr(), nx[1].data_ptr()
        p.nx_relu, p.out_bits = int(nx_relu), (None if bits is None else bits.data_ptr())
        if nx2 is not None:
            p.nx2_scale, p.nx2_shift, p.nx2_relu, p.out_bits2 = nx2[0].data_ptr(), nx2[1].data_ptr(), int(nx2_relu), bits2.data_ptr()
        p.out = out.data_ptr()
        rc = lib.bnn_stem_tc_run(ctypes.byref(p), flags, _stream_ptr(dev))
    native.check(rc, "bnn_stem_tc_run")
    pk = None if bits is None else PackedActivations(bits, n, 64, ho, wo)
    if nx2 is not None:
        return out, pk, PackedActivations(bits2, n, 64, ho, wo)
    return out, pk


def ubench(which: int, iters: int = 200) -> float:
    """Integer-pipe micro-benchmark (giga warp-lane operations / s), see include/bnn_b200.h."""
    v = ctypes.c_double(0.0)
    native.check(native.lib().bnn_ubench(which, iters, ctypes.byref(v)), "bnn_ubench")
    return float(v.value)
