"""TEST INFRASTRUCTURE ONLY -- torch-CPU restatement of the reference's float simulation.

The reference's binarized layer is, line by line (reference bnn/layers/conv.py:90-97,
bnn/ops.py:63-66,116-140,200-202):

    xb  = sign(x)                                  # BasicInputBinarizer
    wc  = W - W.mean(1, keepdim)  (optional)       # XNORWeightBinarizer.center_weights
    a   = ||wc||_1 over (c_in,kh,kw) / n           # XNORWeightBinarizer._compute_alpha
    wb  = sign(wc) * a
    y   = conv2d(xb, wb, bias, stride, padding, dilation)   # torch F.conv2d (third party)
    y   = y * alpha_post                           # BasicScaleBinarizer (in place upstream)

This module restates exactly that with the same torch calls, so on CPU it is bit-identical to
the reference (checked against the real reference by tests/golden/make_golden.py in the build
container).  It provides (1) functional forms used by the parity tests, (2) ``FloatSimConv2d`` /
``FloatSimLinear`` modules and ``mirror_model`` which rebuilds a model prepared with
``bnn_b200`` as its float-simulated twin for whole-model parity and for the CPU baseline that
``bench.py`` times (``cpu_baseline.kind = "port"``).
"""
import copy
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


def binarize_weight(w: torch.Tensor, compute_alpha: bool, center_weights: bool) -> torch.Tensor:
    if center_weights:
        w = w - w.mean(1, keepdim=True)
    s = torch.sign(w)
    if not compute_alpha:
        return s
    n = w[0].nelement()
    alpha = w.norm(1, w.dim() - 1, keepdim=True)
    if w.dim() > 2:
        alpha = alpha.sum(list(range(1, w.dim() - 1)), keepdim=True)
    return s * (alpha / n)


def conv2d(x, weight, bias=None, post=None, stride=1, padding=0, dilation=1, compute_alpha=True,
           center_weights=False, groups=1):
    y = F.conv2d(torch.sign(x), binarize_weight(weight, compute_alpha, center_weights), bias, stride, padding,
                 dilation, groups)
    return y if post is None else y * post.reshape(1, -1, 1, 1)


def conv1d(x, weight, bias=None, post=None, stride=1, padding=0, dilation=1, compute_alpha=True,
           center_weights=False):
    y = F.conv1d(torch.sign(x), binarize_weight(weight, compute_alpha, center_weights), bias, stride, padding,
                 dilation)
    return y if post is None else y * post.reshape(1, -1, 1)


def linear(x, weight, bias=None, post=None, compute_alpha=True, center_weights=False):
    y = F.linear(torch.sign(x), binarize_weight(weight, compute_alpha, center_weights), bias)
    return y if post is None else y * post.reshape(1, -1)


class FloatSimConv2d(nn.Module):
    def __init__(self, weight, bias, post, stride, padding, dilation, compute_alpha, center_weights, groups=1):
        super().__init__()
        self.weight, self.bias, self.post = weight, bias, post
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups
        self.compute_alpha, self.center_weights = compute_alpha, center_weights

    def forward(self, x):
        return conv2d(x, self.weight, self.bias, self.post, self.stride, self.padding, self.dilation,
                      self.compute_alpha, self.center_weights, self.groups)


class FloatSimLinear(nn.Module):
    def __init__(self, weight, bias, post, compute_alpha, center_weights):
        super().__init__()
        self.weight, self.bias, self.post = weight, bias, post
        self.compute_alpha, self.center_weights = compute_alpha, center_weights

    def forward(self, x):
        return linear(x, self.weight, self.bias, self.post, self.compute_alpha, self.center_weights)


def _cpu(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.detach().to("cpu", torch.float32).clone()


def mirror_model(prepared: nn.Module) -> nn.Module:
    """CPU float-simulated twin of a model prepared with bnn_b200: every binarized Conv2d / Linear
    (recognised by its ``weight_pre_process`` attribute) is replaced by a FloatSim module that
    holds a CPU copy of the same parameters; everything else is deep-copied to CPU unchanged."""
    twin = copy.deepcopy(prepared).to("cpu")
    for name, mod in list(twin.named_modules()):
        wpre = getattr(mod, "weight_pre_process", None)
        if wpre is None or not hasattr(wpre, "compute_alpha"):
            continue
        post_mod = getattr(mod, "activation_post_process", None)
        post = _cpu(post_mod.alpha).reshape(-1) if hasattr(post_mod, "alpha") else None
        if isinstance(mod, nn.Conv2d):
            new = FloatSimConv2d(_cpu(mod.weight), _cpu(mod.bias), post, mod.stride, mod.padding, mod.dilation,
                                 bool(wpre.compute_alpha), bool(wpre.center_weights), mod.groups)
        elif isinstance(mod, nn.Linear):
            new = FloatSimLinear(_cpu(mod.weight), _cpu(mod.bias), post, bool(wpre.compute_alpha),
                                 bool(wpre.center_weights))
        else:
            continue
        parent_path, _, attr = name.rpartition(".")
        parent = twin.get_submodule(parent_path) if parent_path else twin
        if parent is twin and not attr:
            return new
        setattr(parent, attr, new)
    return twin.eval()
