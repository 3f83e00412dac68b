"""Process-wide switches of the host side.

``floatsim`` -- the float simulation.  The default forward of every binarized layer in inference is the
CUDA path and nothing else: CPU tensors or a missing library raise ``NativeError``.  Training (straight-through
gradients, reference bnn/ops.py:68-73) needs the reference's fp32 simulation: a layer in ``train()`` mode with
autograd enabled routes to it automatically with a one-time warning (the reference's own behaviour,
bnn/layers/conv.py:90-97); everything else needs the explicit opt-in ``runtime.floatsim(True)`` / the context
manager ``runtime.floatsim_enabled()``.  The simulation is built from torch ops on whatever device the tensors
are on, never from the test oracle.

The switches are process-wide defaults (so the replica threads ``nn.DataParallel`` spawns see what the main
thread set, as the reference's examples/cifar10.py trains under DataParallel) with a per-thread override used by
the context managers.
"""
import contextlib
import threading

_defaults = {"floatsim": False, "flags": 0, "autotune": True, "shortcut_max_cin": 1024}
_local = threading.local()
_UNSET = object()


def _get(name):
    v = getattr(_local, name, _UNSET)
    return _defaults[name] if v is _UNSET else v


def floatsim(enabled: bool = None) -> bool:
    """Get (no argument) or set the process-wide float-simulation opt-in; returns the current value."""
    if enabled is not None:
        _defaults["floatsim"] = bool(enabled)
        if hasattr(_local, "floatsim"):
            del _local.floatsim
    return _get("floatsim")


@contextlib.contextmanager
def floatsim_enabled(enabled: bool = True):
    """Thread-local override of ``floatsim`` for the duration of the block."""
    prev = getattr(_local, "floatsim", _UNSET)
    _local.floatsim = bool(enabled)
    try:
        yield
    finally:
        if prev is _UNSET:
            del _local.floatsim
        else:
            _local.floatsim = prev


def kernel_flags(value: int = None) -> int:
    """Debug flags OR-ed into every conv launch (native.F_STAGE_LDG, native.F_NO_CSA)."""
    if value is not None:
        _defaults["flags"] = int(value)
    return _get("flags")


def autotune(enabled: bool = None) -> bool:
    """Get / set tile-plan autotuning (default on): the first launch of each conv geometry outside a
    CUDA-graph capture times the cost model's best few plans on the device and caches the winner."""
    if enabled is not None:
        _defaults["autotune"] = bool(enabled)
    return _get("autotune")


def shortcut_max_cin(value: int = None) -> int:
    """Largest input-channel count for which the fused engine runs a down-sampling shortcut as ONE kernel
    (``bnn_shortcut_fwd``: pooled planes stay in shared memory); wider shortcuts take pool+pack and the tiled conv."""
    if value is not None:
        _defaults["shortcut_max_cin"] = int(value)
    return _get("shortcut_max_cin")
