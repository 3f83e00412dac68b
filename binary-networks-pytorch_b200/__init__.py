"""bnn_b200 -- B200-native binary-convolution inference behind the ``bnn`` API.

    import bnn_b200 as bnn
    model = bnn.prepare_binary_model(model, bnn.BConfig(...), ignore_layers_name=['_first_', '_last_'])

keeps the surface of 1adrianb/binary-networks-pytorch (``BConfig``, ``prepare_binary_model``,
``layers.Conv2d/Linear``, ``ops`` binarizers; reference bnn/__init__.py:1-4) while every binarized
layer's forward runs hand-written sm_100a kernels through the C ABI in include/bnn_b200.h.
"""
from .bconfig import BConfig, Identity
from . import ops, layers, runtime, native, functional
from .convert import (B200_MODULE_MAPPING, DEFAULT_MODULE_MAPPING, get_modules_to_binarize,
                      get_unique_devices_, invalidate, mapping_for_reference, prepare_binary_model,
                      swap_modules_by_name)

__version__ = "0.1.0"
