"""Binarization config threaded through conversion -> layers -> binarizers.

Mirrors the reference's ``bnn.bconfig`` surface (reference bnn/bconfig.py:6-25):
``BConfig`` holds three *factories* (classes or ``with_args`` partials), never
instances, and ``Identity`` is the two-argument no-op post-process.
"""
from dataclasses import dataclass, fields
from typing import Any, Callable

import torch
import torch.nn as nn


class Identity(nn.Identity):
    """Post-process placeholder: ``f(layer_out, layer_in) -> layer_out`` (reference bnn/bconfig.py:6-8)."""

    def forward(self, layer_out: torch.Tensor, layer_in: torch.Tensor = None) -> torch.Tensor:
        return layer_out


@dataclass
class BConfig:
    activation_pre_process: Callable[..., nn.Module] = nn.Identity
    activation_post_process: Callable[..., nn.Module] = Identity
    weight_pre_process: Callable[..., nn.Module] = nn.Identity

    def __post_init__(self) -> None:
        # reference bnn/bconfig.py:17-25 -- a module *instance* is a usage error
        for f in fields(self):
            if isinstance(getattr(self, f.name), nn.Module):
                raise ValueError("BConfig received an instance, please pass the class instead.")

    def build(self, owner: nn.Module) -> "tuple[Any, Any, Any]":
        """Instantiate (pre, post, weight) for one layer; the post factory gets the layer
        (reference bnn/layers/conv.py:86-88)."""
        return (self.activation_pre_process(), self.activation_post_process(owner), self.weight_pre_process())
