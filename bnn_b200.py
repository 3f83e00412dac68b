"""Import shim: the package directory is named ``binary-networks-pytorch_b200`` (not a valid
Python identifier), so ``import bnn_b200`` resolves to this file, which loads that directory as
the package ``bnn_b200`` and replaces itself in ``sys.modules``."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "binary-networks-pytorch_b200")
_spec = importlib.util.spec_from_file_location(
    "bnn_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_module = importlib.util.module_from_spec(_spec)
sys.modules["bnn_b200"] = _module
_spec.loader.exec_module(_module)
