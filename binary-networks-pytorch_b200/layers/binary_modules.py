"""B200 binarized layers: drop-in for ``bnn.layers.{Conv2d, Conv1d, Linear}``.

Contract mirrored from the reference (bnn/layers/conv.py:65-117, bnn/layers/linear.py:9-44):
subclass of the torch module, constructor takes the torch arguments plus ``bconfig``; the
instance exposes ``bconfig``, ``activation_pre_process``, ``activation_post_process``,
``weight_pre_process``; ``from_module`` re-binds (shares) ``weight``/``bias`` with the source
module; ``state_dict`` keys are ``weight``, ``bias``, ``activation_post_process.alpha``.

What changes is ``forward``: instead of ``post(F.conv2d(sign(x), sign(W)*alpha, bias), x)`` on
dense fp32 tensors it runs   bit-pack(x) -> XNOR/popcount kernel with fused epilogue
through the C ABI (include/bnn_b200.h).  Packed weights are a cache keyed on the weight's
storage and version counter, so ``load_state_dict`` / optimizer steps invalidate it.
"""
import warnings
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as BF
from .. import runtime
from ..bconfig import BConfig
from ..native import NativeError


_warned_training = False


def _warn_training_floatsim() -> None:
    global _warned_training
    if not _warned_training:
        _warned_training = True
        warnings.warn("bnn_b200: a binarized layer in train() mode with autograd enabled runs the reference's fp32 "
                      "simulation (straight-through gradients, torch ops); the B200 kernels are the eval()/no_grad "
                      "forward path.  This warning is shown once.", RuntimeWarning, stacklevel=3)


class NotLowerable(Exception):
    """The layer's configuration has no packed lowering (reason in str(e))."""


@dataclass
class _Lowering:
    center_weights: bool
    compute_alpha: bool
    fused_post: bool          # BasicScaleBinarizer with one scale per output channel
    has_post: bool            # any post-process other than an identity


def _class_named(obj, name: str) -> bool:
    """True for our class *or* the reference's class of that name (``bnn.ops.X``): the modules
    are matched by name so a BConfig built from either package lowers the same way."""
    return any(k.__name__ == name and k.__module__.split(".")[-1] in ("ops", "bconfig")
               for k in type(obj).__mro__)


def _is_identity(mod) -> bool:
    return isinstance(mod, nn.Identity)


class _BinaryLayer:
    """Shared behaviour of the three layer types (mixed in before the torch class)."""

    _FLOAT_MODULE = None
    _ACCEPTED_SOURCES: Tuple[type, ...] = ()

    # -- construction -----------------------------------------------------------------------
    def _attach(self, bconfig: BConfig) -> None:
        assert bconfig, "bconfig is required for a binarized module"
        self.bconfig = bconfig
        self.activation_pre_process = bconfig.activation_pre_process()
        self.activation_post_process = bconfig.activation_post_process(self)
        self.weight_pre_process = bconfig.weight_pre_process()
        self._pack_key = None
        self._packed = None

    @classmethod
    def register_source(cls, foreign_type: type) -> None:
        """Allow ``from_module`` to convert instances of ``foreign_type`` (e.g. the reference's
        own ``bnn.layers.Conv2d`` when a model was prepared with the reference first)."""
        if foreign_type not in cls._ACCEPTED_SOURCES:
            cls._ACCEPTED_SOURCES = cls._ACCEPTED_SOURCES + (foreign_type,)

    @classmethod
    def _resolve_source(cls, mod: nn.Module, bconfig: Optional[BConfig]) -> BConfig:
        ok = type(mod) in (cls._FLOAT_MODULE, cls) + cls._ACCEPTED_SOURCES
        assert ok, "bnn." + cls.__name__ + ".from_float only works for " + cls._FLOAT_MODULE.__name__
        if not bconfig:
            assert hasattr(mod, "bconfig"), "The input modele requires a predifined bconfig"
            assert mod.bconfig, "The input modele bconfig is invalid"
            bconfig = mod.bconfig
        return bconfig

    @staticmethod
    def _adopt(new: nn.Module, mod: nn.Module, bconfig: BConfig, update: bool) -> nn.Module:
        new.weight = mod.weight       # shared storage, reference bnn/layers/conv.py:111-112
        new.bias = mod.bias
        if update:
            # carry over learned binarizer parameters of matching shape (reference helpers.py:7-17)
            for slot in ("activation_pre_process", "activation_post_process", "weight_pre_process"):
                src, dst = getattr(mod, slot, None), getattr(new, slot, None)
                if src is None or dst is None:
                    continue
                dst_params = dict(dst.named_parameters())
                for pname, p in src.named_parameters():
                    q = dst_params.get(pname)
                    if q is not None and q.shape == p.shape:
                        q.data.copy_(p.data)
        return new

    # -- lowering analysis --------------------------------------------------------------------
    def _is_float_layer(self) -> bool:
        """A layer configured with identity binarizers is a plain fp32 layer (the reference's
        'skip binarization' recipe, test/test_binarize.py:74-93); it is not on the binary path."""
        return _is_identity(self.activation_pre_process) or _is_identity(self.weight_pre_process)

    def _lowering(self) -> _Lowering:
        pre, wpre, post = self.activation_pre_process, self.weight_pre_process, self.activation_post_process
        if _class_named(pre, "AdvancedInputBinarizer"):
            if getattr(pre, "derivative_funct", None) is not torch.tanh or not (pre.t > 0):
                raise NotLowerable("AdvancedInputBinarizer with a non-default surrogate")
        elif not _class_named(pre, "BasicInputBinarizer"):
            raise NotLowerable(f"activation_pre_process {type(pre).__name__} has no packed lowering")
        if not _class_named(wpre, "XNORWeightBinarizer"):
            raise NotLowerable(f"weight_pre_process {type(wpre).__name__} has no packed lowering")
        fused = False
        has_post = not _is_identity(post)
        if has_post and _class_named(post, "BasicScaleBinarizer"):
            a = post.alpha
            fused = a.numel() == self._out_channels() and a.dim() >= 2 and a.shape[1] == a.numel()
        self._check_geometry()
        return _Lowering(bool(wpre.center_weights), bool(wpre.compute_alpha), fused, has_post)

    def _check_geometry(self) -> None:
        pass

    def _out_channels(self) -> int:
        return self.weight.shape[0]

    # -- packed weight cache ------------------------------------------------------------------
    def _packed_weights(self, low: _Lowering) -> BF.PackedWeights:
        w = self.weight
        key = (w.data_ptr(), w._version, w.device, tuple(w.shape), low.center_weights, low.compute_alpha)
        if self._pack_key != key:
            # exactly-zero (centred) weights have sign 0 in the reference (bnn/ops.py:66,136): pack_weights then returns a
            # ternary pair of planes and the layer runs as two binary launches + an exact integer mean
            self._packed, self._pack_key = BF.pack_weights(w, low.center_weights, low.compute_alpha), key
        return self._packed

    def repack(self) -> None:
        """Drop the packed-weight cache.  The cache follows the weight's storage pointer and version counter, so
        ``load_state_dict``, optimizer steps and in-place tensor ops invalidate it automatically; writes made THROUGH
        ``weight.data`` (``w.data.clamp_(-1, 1)``, ``w.data.copy_(...)``) do not bump the version counter -- call
        ``repack()`` (or ``bnn_b200.invalidate(model)``) after such updates."""
        self._pack_key = self._packed = None

    # -- forward ------------------------------------------------------------------------------
    def _wants_autograd(self, x: torch.Tensor) -> bool:
        if not torch.is_grad_enabled() or not self.training:
            return False
        return x.requires_grad or any(p.requires_grad for p in self.parameters())

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if self._is_float_layer():
            return self._forward_floatsim(input)
        reason = None
        try:
            low = self._lowering()
        except NotLowerable as e:
            low, reason = None, str(e)
        if low is not None:
            if self._wants_autograd(input):
                # train() + autograd: the reference's own forward (bnn/layers/conv.py:90-97) with the straight-through
                # estimator (bnn/ops.py:68-73), built from torch ops; the packed kernels are inference-only
                _warn_training_floatsim()
                return self._forward_floatsim(input)
            elif not input.is_cuda:
                reason = f"input is on {input.device}; the B200 path has no CPU implementation"
            elif input.dtype != torch.float32:
                reason = f"input dtype {input.dtype} (expected float32)"
            else:
                out = self._forward_packed(input, low)
                if low.has_post and not low.fused_post:
                    out = self.activation_post_process(out, input)
                return out
        if runtime.floatsim():
            return self._forward_floatsim(input)
        raise NativeError(
            f"{type(self).__name__}: cannot run the CUDA path ({reason}). There is no silent fallback; "
            "enable the fp32 simulation explicitly with bnn_b200.runtime.floatsim(True) if that is intended.")

    def _post_scale(self, low: _Lowering) -> Optional[torch.Tensor]:
        if not low.fused_post:
            return None
        return self.activation_post_process.alpha.detach().reshape(-1)

    def _bias(self) -> Optional[torch.Tensor]:
        return None if self.bias is None else self.bias.detach()


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class _ConvNd(_BinaryLayer):
    def _check_geometry(self) -> None:
        if self.padding_mode != "zeros":
            raise NotLowerable(f"padding_mode={self.padding_mode}")
        self._resolved_padding()

    def _resolved_padding(self):
        if isinstance(self.padding, str):
            if self.padding == "valid":
                return tuple(0 for _ in self.kernel_size)
            total = [d * (k - 1) for d, k in zip(self.dilation, self.kernel_size)]
            if any(t % 2 for t in total):
                raise NotLowerable("asymmetric 'same' padding")
            return tuple(t // 2 for t in total)
        return tuple(self.padding)

    @classmethod
    def from_module(cls, mod: nn.Module, bconfig: BConfig = None, update: bool = False):
        bconfig = cls._resolve_source(mod, bconfig)
        new = cls(mod.in_channels, mod.out_channels, mod.kernel_size, stride=mod.stride, padding=mod.padding,
                  dilation=mod.dilation, groups=mod.groups, bias=mod.bias is not None,
                  padding_mode=mod.padding_mode, bconfig=bconfig)
        return cls._adopt(new, mod, bconfig, update)

    def _forward_floatsim(self, input: torch.Tensor) -> torch.Tensor:
        x = self.activation_pre_process(input)
        y = self._conv_forward(x, self.weight_pre_process(self.weight), self.bias)
        return self.activation_post_process(y, input)


class Conv2d(_ConvNd, nn.Conv2d):
    """Binarized 2-D convolution (reference bnn/layers/conv.py:65-117)."""
    _FLOAT_MODULE = nn.Conv2d

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, padding_mode="zeros", bconfig: BConfig = None) -> None:
        nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                           dilation=dilation, groups=groups, bias=bias, padding_mode=padding_mode)
        self._attach(bconfig)

    def _forward_packed(self, x: torch.Tensor, low: _Lowering) -> torch.Tensor:
        if x.dim() != 4:
            raise NativeError(f"Conv2d expects a 4-D input, got {tuple(x.shape)}")
        if self.groups == 1:
            wts = self._packed_weights(low)
            act = BF.pack_activations(x)
            return BF.bconv2d(act, wts, self._bias(), self._post_scale(low), _pair(self.stride),
                              self._resolved_padding(), _pair(self.dilation), use_alpha=low.compute_alpha,
                              flags=runtime.kernel_flags())
        return self._forward_grouped(x, low)

    def _forward_grouped(self, x: torch.Tensor, low: _Lowering) -> torch.Tensor:
        """groups > 1 (e.g. the BATS cells, bnn/models/layers/bats_ops.py:41-52): one launch pair per group on
        strided channel slices of the input and of the output -- the kernels take element strides, so no copies."""
        g = self.groups
        cin_g, cout_g = self.in_channels // g, self.out_channels // g
        packed = self._packed_group_weights(low)
        kh, kw = self.kernel_size
        sh, sw = _pair(self.stride)
        ph, pw = self._resolved_padding()
        dh, dw = _pair(self.dilation)
        ho = (x.shape[2] + 2 * ph - dh * (kh - 1) - 1) // sh + 1
        wo = (x.shape[3] + 2 * pw - dw * (kw - 1) - 1) // sw + 1
        out = torch.empty((x.shape[0], self.out_channels, ho, wo), dtype=torch.float32, device=x.device)
        bias, post = self._bias(), self._post_scale(low)
        for i in range(g):
            act = BF.pack_activations(x[:, i * cin_g:(i + 1) * cin_g])
            sl = slice(i * cout_g, (i + 1) * cout_g)
            BF.bconv2d(act, packed[i], None if bias is None else bias[sl], None if post is None else post[sl],
                       (sh, sw), (ph, pw), (dh, dw), use_alpha=low.compute_alpha, flags=runtime.kernel_flags(),
                       out=out[:, sl])
        return out

    def _packed_group_weights(self, low: _Lowering):
        w = self.weight
        key = ("groups", w.data_ptr(), w._version, w.device, tuple(w.shape), low.center_weights, low.compute_alpha)
        if self._pack_key != key:
            cout_g = self.out_channels // self.groups
            packed = [BF.pack_weights(w[i * cout_g:(i + 1) * cout_g], low.center_weights, low.compute_alpha)
                      for i in range(self.groups)]
            self._packed, self._pack_key = packed, key
        return self._packed


class Conv1d(_ConvNd, nn.Conv1d):
    """Binarized 1-D convolution (reference bnn/layers/conv.py:10-62), run as a 1 x k Conv2d."""
    _FLOAT_MODULE = nn.Conv1d

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, padding_mode="zeros", bconfig: BConfig = None) -> None:
        nn.Conv1d.__init__(self, in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                           dilation=dilation, groups=groups, bias=bias, padding_mode=padding_mode)
        self._attach(bconfig)

    def _check_geometry(self) -> None:
        if self.groups != 1:
            raise NotLowerable(f"groups={self.groups}")
        super()._check_geometry()

    def _forward_packed(self, x: torch.Tensor, low: _Lowering) -> torch.Tensor:
        if x.dim() != 3:
            raise NativeError(f"Conv1d expects a 3-D input, got {tuple(x.shape)}")
        wts = self._packed_weights(low)
        act = BF.pack_activations(x.unsqueeze(2))
        y = BF.bconv2d(act, wts, self._bias(), self._post_scale(low), (1, self.stride[0]),
                       (0, self._resolved_padding()[0]), (1, self.dilation[0]), use_alpha=low.compute_alpha,
                       flags=runtime.kernel_flags())
        return y.squeeze(2)


class Linear(_BinaryLayer, nn.Linear):
    """Binarized fully-connected layer (reference bnn/layers/linear.py:9-44)."""
    _FLOAT_MODULE = nn.Linear

    def __init__(self, in_features: int, out_features: int, bias: bool = True, bconfig: BConfig = None) -> None:
        nn.Linear.__init__(self, in_features, out_features, bias)
        self._attach(bconfig)

    @classmethod
    def from_module(cls, mod: nn.Module, bconfig: BConfig = None, update: bool = False):
        bconfig = cls._resolve_source(mod, bconfig)
        new = cls(mod.in_features, mod.out_features, bias=mod.bias is not None, bconfig=bconfig)
        return cls._adopt(new, mod, bconfig, update)

    def _forward_floatsim(self, input: torch.Tensor) -> torch.Tensor:
        x = self.activation_pre_process(input)
        return self.activation_post_process(F.linear(x, self.weight_pre_process(self.weight), self.bias), input)

    def _forward_packed(self, x: torch.Tensor, low: _Lowering) -> torch.Tensor:
        wts = self._packed_weights(low)
        rows = x.reshape(-1, self.in_features)
        act = BF.pack_activations(rows, linear_rows=True)
        y = BF.blinear(act, wts, self._bias(), self._post_scale(low), use_alpha=low.compute_alpha,
                       flags=runtime.kernel_flags())
        return y.reshape(*x.shape[:-1], self.out_features)
