"""Staged-binarization recipe driver: same surface as the reference's ``bnn.engine.BinaryChef``
(reference bnn/engine.py:23-79): a YAML file lists steps, each naming the activation pre-process,
activation post-process and weight pre-process (plus optional ``args`` and ``ignore_layer_names``);
``chef.next(model)`` applies the next step with ``prepare_binary_model``.

Differences in mechanism, not behaviour: names are resolved in a registry (``bnn_b200.ops``, ``nn.*``,
``Identity`` and user-supplied classes) instead of ``eval``-ing strings, ``easydict`` is not needed, and
the keys ``name`` / ``args`` are matched case-insensitively (upstream's shipped xnor-net.yaml spells
``NAME`` and cannot be loaded by upstream itself).
"""
from typing import Any, Callable, Dict, List

import torch
import torch.nn as nn
import yaml

from . import ops
from .bconfig import BConfig, Identity
from .convert import prepare_binary_model


def _lower_keys(d: Dict[str, Any]) -> Dict[str, Any]:
    return {str(k).lower(): v for k, v in d.items()}


class BinaryChef:
    def __init__(self, config: str, user_modules: List[Callable[..., nn.Module]] = []) -> None:
        with open(config) as fh:
            raw = yaml.safe_load(fh)
        self.config = [_lower_keys(raw[k]) for k in raw.keys()]
        self.current_step = 0
        self._registry: Dict[str, Any] = {name: getattr(ops, name) for name in ops.__all__}
        self._registry["Identity"] = Identity
        for mod in user_modules:
            self._registry[mod.__name__] = mod

    def __len__(self) -> int:
        return len(self.config)

    def get_num_steps(self) -> int:
        return len(self)

    def _resolve(self, name: str) -> Any:
        if name in self._registry:
            return self._registry[name]
        for prefix, namespace in (("nn.", nn), ("torch.nn.", nn), ("torch.", torch)):
            if name.startswith(prefix) and hasattr(namespace, name[len(prefix):]):
                return getattr(namespace, name[len(prefix):])
        raise NameError(f"BinaryChef: unknown binarizer {name!r}")

    def _value(self, v: Any) -> Any:
        if isinstance(v, str):
            try:
                return self._resolve(v)          # e.g. derivative_funct: torch.tanh
            except NameError:
                return v
        return v

    def _factory(self, spec: Dict[str, Any]) -> Any:
        spec = _lower_keys(spec)
        target = self._resolve(spec["name"])
        args = spec.get("args")
        if args:
            return target.with_args(**{k: self._value(v) for k, v in args.items()})
        return target

    def run_step(self, model: nn.Module, step: int) -> nn.Module:
        assert len(self) > step
        cfg = self.config[step]
        bconfig = BConfig(activation_pre_process=self._factory(cfg["pre_activation"]),
                          activation_post_process=self._factory(cfg["post_activation"]),
                          weight_pre_process=self._factory(cfg["weight"]))
        return prepare_binary_model(model, bconfig=bconfig, ignore_layers_name=cfg.get("ignore_layer_names", []))

    def next(self, model: nn.Module) -> nn.Module:
        self.current_step += 1
        return self.run_step(model, self.current_step - 1)
