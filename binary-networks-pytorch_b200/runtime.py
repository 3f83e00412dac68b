"""Process-wide switches of the host side.

``floatsim`` -- opt-in float simulation.  The default forward of every binarized layer is the
CUDA path and nothing else: CPU tensors, a missing library or a training-mode autograd request
raise ``NativeError``.  Training (straight-through gradients) needs the reference's fp32
simulation; it is enabled explicitly with ``runtime.floatsim(True)`` or the context manager
``runtime.floatsim_enabled()`` and is built from torch ops on whatever device the tensors are on.
"""
import contextlib
import threading

_state = threading.local()


def floatsim(enabled: bool = None) -> bool:
    """Get (no argument) or set the float-simulation opt-in; returns the current value."""
    if enabled is not None:
        _state.floatsim = bool(enabled)
    return getattr(_state, "floatsim", False)


@contextlib.contextmanager
def floatsim_enabled(enabled: bool = True):
    prev = floatsim()
    floatsim(enabled)
    try:
        yield
    finally:
        floatsim(prev)


def kernel_flags(value: int = None) -> int:
    """Debug flags OR-ed into every conv launch (native.F_STAGE_LDG, native.F_NO_CSA)."""
    if value is not None:
        _state.flags = int(value)
    return getattr(_state, "flags", 0)


def autotune(enabled: bool = None) -> bool:
    """Get / set tile-plan autotuning (default on): the first launch of each conv geometry outside a
    CUDA-graph capture times the cost model's best few plans on the device and caches the winner."""
    if enabled is not None:
        _state.autotune = bool(enabled)
    return getattr(_state, "autotune", True)
