"""Binarizers of the reference's ``bnn.ops`` surface (reference bnn/ops.py:10-221).

On the B200 path these modules are *descriptors*: ``layers`` inspects them (class +
hyper-parameters) and lowers ``BasicInputBinarizer`` to the activation bit-pack kernel,
``XNORWeightBinarizer`` to the prepare-time weight pack, and ``BasicScaleBinarizer`` to the
fused epilogue scale.  Their ``forward`` methods keep the reference's fp32 semantics (with the
straight-through backward) for the explicit float-simulation mode used in training.
"""
import math
from typing import Any, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Factory:
    """``cls.with_args(**kw)`` -> zero-state factory; chaining merges kwargs, later wins
    (same observable behaviour as the reference's partial wrapper, bnn/ops.py:10-35)."""

    def __init__(self, target: Any, kwargs: dict) -> None:
        self._target, self._kwargs = target, dict(kwargs)

    def __call__(self, *args: Any, **kwargs: Any) -> Any:
        return self._target(*args, **{**self._kwargs, **kwargs})

    def with_args(self, **kwargs: Any) -> "_Factory":
        return _Factory(self._target, {**self._kwargs, **kwargs})

    def __repr__(self) -> str:
        inner = ", ".join(f"{k}={v!r}" for k, v in self._kwargs.items())
        return f"{getattr(self._target, '__name__', self._target)}.with_args({inner})"


class BinarizerBase(nn.Module):
    """Common base: gives every binarizer the ``with_args`` class factory (bnn/ops.py:40-48)."""

    @classmethod
    def with_args(cls, **kwargs: Any) -> _Factory:
        return _Factory(cls, kwargs)

    def forward(self, *args: Any, **kwargs: Any) -> torch.Tensor:  # pragma: no cover - abstract
        raise NotImplementedError


class SignActivation(torch.autograd.Function):
    """sign with hard-tanh straight-through gradient (bnn/ops.py:51-73). sign(+-0) = 0."""

    @staticmethod
    def forward(ctx, x: torch.Tensor) -> torch.Tensor:
        ctx.save_for_backward(x)
        return torch.sign(x)

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor) -> torch.Tensor:
        (x,) = ctx.saved_tensors
        return grad_out.masked_fill(x.abs() >= 1, 0)


class SignActivationStochastic(SignActivation):
    """Stochastic binarization (bnn/ops.py:76-92); training-only, mutates its input like upstream."""

    @staticmethod
    def forward(ctx, x: torch.Tensor) -> torch.Tensor:
        ctx.save_for_backward(x)
        noise = torch.rand_like(x) - 0.5
        return x.add_(1).div_(2).add_(noise).clamp_(0, 1).round_().mul_(2).sub_(1)


class BasicInputBinarizer(BinarizerBase):
    """Activation binarizer, ``sign(x)`` (bnn/ops.py:143-152) -> B200: ``bnn_pack_act_f32``."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return SignActivation.apply(x)


class StochasticInputBinarizer(BinarizerBase):
    """bnn/ops.py:155-164; no packed lowering (RNG, training-only)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return SignActivationStochastic.apply(x)


class AdvancedInputBinarizer(BinarizerBase):
    """sign forward with a smooth surrogate gradient (bnn/ops.py:167-177)."""

    def __init__(self, derivative_funct=torch.tanh, t: int = 5) -> None:
        super().__init__()
        self.derivative_funct, self.t = derivative_funct, t

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.derivative_funct(x * self.t)
        with torch.no_grad():
            y = torch.sign(y)
        return y


class XNORWeightBinarizer(BinarizerBase):
    """XNOR-Net weight binarizer (bnn/ops.py:95-140) -> B200: ``bnn_pack_weight_f32``.

    ``center_weights`` subtracts the mean over the input-channel axis; ``compute_alpha`` scales
    ``sign(w)`` by the mean absolute (centred) weight of each output channel."""

    def __init__(self, compute_alpha: bool = True, center_weights: bool = False) -> None:
        super().__init__()
        self.compute_alpha, self.center_weights = compute_alpha, center_weights

    @staticmethod
    def _compute_alpha(w: torch.Tensor) -> torch.Tensor:
        if w.dim() not in (2, 3, 4):
            raise ValueError(f"Expected ndims equal with 2 or 4, but found {w.dim()}")
        # L1 norm along the last axis, summed over the remaining non-output axes (ops.py:117-123)
        alpha = w.norm(1, w.dim() - 1, keepdim=True)
        if w.dim() > 2:
            alpha = alpha.sum(list(range(1, w.dim() - 1)), keepdim=True)
        return alpha.div_(w[0].nelement())

    def forward(self, w: torch.Tensor) -> torch.Tensor:
        if self.center_weights:
            w = w - w.mean(1, keepdim=True)
        if not self.compute_alpha:
            return SignActivation.apply(w)
        return SignActivation.apply(w).mul_(self._compute_alpha(w).expand_as(w))

    def extra_repr(self) -> str:
        return f"compute_alpha={self.compute_alpha}, center_weights={self.center_weights}"


class BasicScaleBinarizer(BinarizerBase):
    """Learned per-output-channel scale applied in place to the layer output
    (bnn/ops.py:180-205) -> B200: ``post[]`` of the fused epilogue."""

    def __init__(self, module: nn.Module, shape: Optional[List[int]] = None) -> None:
        super().__init__()
        if isinstance(module, nn.Linear):
            channels = module.out_features
        elif hasattr(module, "out_channels"):
            channels = module.out_channels
        else:
            raise Exception("Unknown layer of type {} missing out_channels".format(type(module)))
        if shape is None:
            shape = [1, channels] + [1] * (module.weight.dim() - 2)
        self.alpha = nn.Parameter(torch.ones(*shape))

    def forward(self, layer_out: torch.Tensor, layer_in: torch.Tensor = None) -> torch.Tensor:
        return layer_out.mul_(self.alpha)

    def extra_repr(self) -> str:
        return "{}".format(list(self.alpha.size()))


class XNORScaleBinarizer(BinarizerBase):
    """Input-dependent XNOR-Net scale K = box-filter(mean_c |x|) (bnn/ops.py:208-221).

    Upstream's version cannot run (wrong ``super`` and a ``torch.mean`` without input); this one
    implements what it describes.  No packed lowering: it is applied to the kernel's output."""

    def __init__(self, module: nn.Module) -> None:
        super().__init__()
        self.stride, self.padding = module.stride, module.padding
        k = tuple(module.kernel_size)
        self.register_buffer("fixed_weight", torch.ones(1, 1, *k).div_(math.prod(k)), persistent=False)

    def forward(self, layer_out: torch.Tensor, layer_in: torch.Tensor) -> torch.Tensor:
        scale = layer_in.abs().mean(dim=1, keepdim=True)
        scale = F.conv2d(scale, self.fixed_weight, stride=self.stride, padding=self.padding)
        return layer_out.mul_(scale)


__all__ = [
    "BinarizerBase", "SignActivation", "SignActivationStochastic", "XNORWeightBinarizer",
    "BasicInputBinarizer", "StochasticInputBinarizer", "AdvancedInputBinarizer",
    "BasicScaleBinarizer", "XNORScaleBinarizer",
]
