"""Model conversion: swap float layers for the B200 binarized layers.

Same call surface and observable behaviour as the reference's ``bnn.binarize``
(reference bnn/binarize.py:12-141): ``prepare_binary_model(model, bconfig, modules_mapping,
custom_config_layers_name, ignore_layers_name)`` mutates the model in place (or returns the
replacement when the model *is* a single leaf layer).  The only difference is what the default
mapping points at: the CUDA-backed layers of ``bnn_b200.layers``.
"""
import copy
import logging
import re
from dataclasses import asdict
from typing import Dict, List, Optional, Set

import torch
import torch.nn as nn

from . import layers
from .bconfig import BConfig

# float type -> binarized type, plus identity entries so an already converted model can be
# re-configured in place (reference binarize.py:12-18; BinaryChef relies on it)
B200_MODULE_MAPPING: Dict[type, type] = {
    nn.Linear: layers.Linear, nn.Conv2d: layers.Conv2d, nn.Conv1d: layers.Conv1d,
    layers.Linear: layers.Linear, layers.Conv2d: layers.Conv2d, layers.Conv1d: layers.Conv1d,
}
DEFAULT_MODULE_MAPPING = B200_MODULE_MAPPING


def mapping_for_reference(ref_bnn) -> Dict[type, type]:
    """Mapping to hand to the *reference's own* ``bnn.prepare_binary_model(modules_mapping=...)``:
    float layers and the reference's fake-binary layers both convert to the CUDA layers."""
    mapping = dict(B200_MODULE_MAPPING)
    for name in ("Linear", "Conv2d", "Conv1d"):
        ours, theirs = getattr(layers, name), getattr(ref_bnn.layers, name)
        ours.register_source(theirs)
        mapping[theirs] = ours
    return mapping


def _convertible_names(model: nn.Module, mapping: Dict[type, type]) -> List[str]:
    return [name for name, m in model.named_modules() if type(m) in mapping]


def _get_first_layer(model: nn.Module) -> List[str]:
    return _convertible_names(model, DEFAULT_MODULE_MAPPING)[:1]


def _get_last_layer(model: nn.Module) -> List[str]:
    return _convertible_names(model, DEFAULT_MODULE_MAPPING)[-1:]


def _regex_match(model: nn.Module, pattern: str, modules_mapping: Dict[type, type]) -> List[str]:
    rx = re.compile(pattern[1:-1])  # strip the enclosing '$'
    return [name for name in _convertible_names(model, modules_mapping) if rx.search(name) is not None]


# NOTE: upstream binds the two words crosswise ('_last_' resolves to the first convertible layer
# and '_first_' to the last one, reference binarize.py:47-50).  Recipes always pass both, so the
# effect is the same; it is reproduced here so that passing only one behaves like upstream.
_KNOWN_SPECIAL_WORDS = {"_last_": _get_first_layer, "_first_": _get_last_layer}


def get_unique_devices_(module: nn.Module) -> Set[torch.device]:
    return {t.device for t in module.parameters()} | {t.device for t in module.buffers()}


def _expand_ignored(model: nn.Module, names: List[str], mapping: Dict[type, type]) -> List[str]:
    out: List[str] = []
    for name in names:
        if name in _KNOWN_SPECIAL_WORDS:
            out += _KNOWN_SPECIAL_WORDS[name](model)
        elif len(name) >= 2 and name[0] == "$" and name[-1] == "$":
            out += _regex_match(model, name, mapping)
        else:
            out.append(name)
    return out


def get_modules_to_binarize(model: nn.Module, bconfig: BConfig,
                            modules_mapping: Optional[Dict[type, type]] = None,
                            custom_config_layers_name: Dict[str, BConfig] = {},
                            ignore_layers_name: List[str] = []) -> Dict[str, nn.Module]:
    """name -> freshly built binarized module for every convertible, non-ignored layer."""
    mapping = DEFAULT_MODULE_MAPPING if modules_mapping is None else modules_mapping
    ignored = _expand_ignored(model, ignore_layers_name, mapping)
    replacements: Dict[str, nn.Module] = {}
    for name, module in model.named_modules():
        if type(module) not in mapping:
            if name in custom_config_layers_name:
                logging.warning("Module named {} defined in the configuration was not found.".format(name))
            continue
        if name in ignored:
            continue
        cfg = copy.copy(bconfig)
        if name in custom_config_layers_name:     # a per-layer config overrides all three fields
            for field, value in asdict(custom_config_layers_name[name]).items():
                setattr(cfg, field, value)
        devices = get_unique_devices_(module)
        assert len(devices) <= 1, (
            "swap_module only works with cpu or single-device CUDA modules, but got devices {}".format(devices))
        new = mapping[type(module)].from_module(module, cfg)
        if devices:
            new.to(next(iter(devices)))
        replacements[name] = new
    return replacements


def swap_modules_by_name(model: nn.Module, modules_to_replace: Dict[str, nn.Module],
                         modules_mapping: Optional[Dict[type, type]] = None) -> nn.Module:
    """Install the replacements (consumes ``modules_to_replace``); returns the model, or the
    replacement itself when the model is a single convertible leaf (reference binarize.py:121-123)."""
    mapping = DEFAULT_MODULE_MAPPING if modules_mapping is None else modules_mapping
    if next(model.named_children(), None) is None:
        if type(model) in mapping and len(modules_to_replace) == 1:
            return next(iter(modules_to_replace.values()))
        return model
    for name in list(modules_to_replace):
        parent_path, _, attr = name.rpartition(".")
        try:
            parent = model.get_submodule(parent_path) if parent_path else model
        except AttributeError:
            continue
        if attr and type(getattr(parent, attr, None)) in mapping:
            setattr(parent, attr, modules_to_replace.pop(name))
    return model


def prepare_binary_model(model: nn.Module, bconfig: BConfig,
                         modules_mapping: Optional[Dict[type, type]] = None,
                         custom_config_layers_name: Dict[str, BConfig] = {},
                         ignore_layers_name: List[str] = []) -> nn.Module:
    replacements = get_modules_to_binarize(model, bconfig, modules_mapping, custom_config_layers_name,
                                           ignore_layers_name)
    return swap_modules_by_name(model, replacements, modules_mapping)


def invalidate(model: nn.Module) -> nn.Module:
    """Drop every cached derivative of the model's parameters (packed weight planes, alpha, folded BatchNorm
    constants, stem operands).  The caches follow the tensors' version counters, so they only go stale after writes
    made through ``.data`` (``w.data.clamp_(-1, 1)`` in BNN training loops), which do not bump the counter; call this
    after such updates.  Works on a prepared model and on the engines of ``bnn_b200.fuse``.  Returns ``model``."""
    for m in model.modules():
        if hasattr(m, "repack"):
            m.repack()
        if hasattr(m, "invalidate_caches"):
            m.invalidate_caches()
    return model
