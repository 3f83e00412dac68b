"""ctypes binding of the C-ABI library ``csrc/libbnn_b200.so`` (see include/bnn_b200.h).

There is no fallback: if the library is missing or a call fails, ``NativeError`` is raised.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
# BNN_B200_LIB: load another build of the same library (kernel A/B experiments, scripts/build_variant.sh)
LIB_PATH = os.environ.get("BNN_B200_LIB") or os.path.join(_HERE, "csrc", "libbnn_b200.so")

F_STAGE_LDG = 1
F_NO_CSA = 2
Q_ABI_VERSION, Q_SM_ARCH, Q_DEVICE_SMS, Q_LAUNCH_COUNT = 0, 1, 2, 3


class NativeError(RuntimeError):
    pass


class ConvGeom(ctypes.Structure):
    """``struct bnn_conv_geom`` (include/bnn_b200.h)."""
    _fields_ = [(k, c_int32) for k in (
        "n", "c_in", "h", "w", "c_out", "kh", "kw", "stride_h", "stride_w", "pad_h", "pad_w", "dil_h", "dil_w")]


class Epilogue(ctypes.Structure):
    """``struct bnn_epilogue`` (include/bnn_b200.h)."""
    _fields_ = [("scale", c_void_p), ("bias", c_void_p), ("post", c_void_p), ("bn_scale", c_void_p),
                ("bn_shift", c_void_p), ("residual", c_void_p), ("rstride_n", c_int64), ("rstride_c", c_int64),
                ("rstride_h", c_int64), ("rstride_w", c_int64), ("residual_after_act", c_int32), ("act", c_int32),
                ("act_slope", c_void_p), ("out", c_void_p), ("ostride_n", c_int64), ("ostride_c", c_int64),
                ("ostride_h", c_int64), ("ostride_w", c_int64), ("out_bits", c_void_p), ("nx_scale", c_void_p),
                ("nx_shift", c_void_p), ("nx_relu", c_int32), ("bits_before_residual", c_int32)]


class StemTcParams(ctypes.Structure):
    """``struct bnn_stem_tc_params`` (include/bnn_b200.h)."""
    _fields_ = [("x", c_void_p), ("x_dtype", c_int32), ("n", c_int32), ("h", c_int32), ("w", c_int32), ("w_ops", c_void_p),
                ("w_log2_scale", c_int32), ("x_log2_scale", c_int32), ("x_amax", c_void_p), ("u8_mean", c_float * 3),
                ("u8_istd", c_float * 3), ("bn_scale", c_void_p), ("bn_shift", c_void_p), ("pool", c_int32),
                ("nx_scale", c_void_p), ("nx_shift", c_void_p), ("nx_relu", c_int32), ("out_bits", c_void_p),
                ("nx2_scale", c_void_p), ("nx2_shift", c_void_p), ("nx2_relu", c_int32), ("out_bits2", c_void_p),
                ("out", c_void_p)]


ACT_NONE, ACT_RELU, ACT_PRELU = 0, 1, 2

_SIGNATURES = {
    "bnn_query": (c_int, [c_int, POINTER(c_int64)]),
    "bnn_strerror": (c_char_p, [c_int]),
    "bnn_act_bits_bytes": (c_size_t, [c_int32] * 4),
    "bnn_weight_bits_bytes": (c_size_t, [c_int32] * 4),
    "bnn_pack_act_f32": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32,
                                 c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "bnn_avgpool_pack_f32": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32,
                                     c_int32, c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "bnn_avgpool2_pack_cl_f32": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_void_p,
                                         c_void_p, c_void_p]),
    "bnn_pack_weight_f32": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "bnn_pack_weight_ternary_f32": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "bnn_conv_split": (c_int, [POINTER(ConvGeom), c_uint32, POINTER(c_int32), POINTER(c_int32)]),
    "bnn_bconv2d_partial_fwd": (c_int, [c_void_p, c_void_p, POINTER(ConvGeom), c_int32, c_int32, c_void_p, c_uint32, c_void_p]),
    "bnn_dot_finish_f32": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                   c_int64, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "bnn_bconv2d_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int64, c_int64, c_int64, c_int64, POINTER(ConvGeom), c_uint32, c_void_p]),
    "bnn_bconv2d_fused_fwd": (c_int, [c_void_p, c_void_p, POINTER(ConvGeom), POINTER(Epilogue), c_uint32, c_void_p]),
    "bnn_blinear_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int32, c_int32, c_int32, c_uint32, c_void_p]),
    "bnn_ubench": (c_int, [c_int32, c_int32, POINTER(c_double)]),
    "bnn_stem_out_hw": (c_int, [c_int32, c_int32, POINTER(c_int32), POINTER(c_int32)]),
    "bnn_stem_fwd": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_uint32, c_void_p]),
    "bnn_stem_mma_weight_bytes": (c_size_t, []),
    "bnn_stem_mma_pack_weight": (c_int, [c_void_p, c_int32, c_void_p, c_void_p]),
    "bnn_stem_mma_fwd": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p]),
    "bnn_amax_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "bnn_stem_tc_weight_bytes": (c_size_t, []),
    "bnn_stem_tc_pack_weight": (c_int, [c_void_p, c_int32, c_void_p, c_void_p]),
    "bnn_stem_tc_run": (c_int, [POINTER(StemTcParams), c_uint32, c_void_p]),
    "bnn_stem_tc_fwd": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p]),
    "bnn_shortcut_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                 c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint32,
                                 c_void_p]),
    "bnn_bconv2d_tune": (c_int, [c_void_p, c_void_p, POINTER(ConvGeom), POINTER(Epilogue), c_uint32, c_int32, c_void_p]),
    "bnn_conv_plan": (c_int, [POINTER(ConvGeom), c_uint32, c_int32, POINTER(c_int32)]),
    "bnn_conv_plan_list": (c_int, [POINTER(ConvGeom), c_uint32, POINTER(c_int32), c_int32, POINTER(c_int32)]),
    "bnn_bconv2d_fused_fwd_plan": (c_int, [c_void_p, c_void_p, POINTER(ConvGeom), POINTER(Epilogue), c_uint32, c_int32,
                                           c_int32, c_int32, c_int32, c_void_p]),
    "bnn_conv_instance": (c_int, [POINTER(ConvGeom), POINTER(Epilogue), c_uint32, c_int32, c_int32, c_int32, c_int32,
                                  POINTER(c_int32)]),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def lib() -> ctypes.CDLL:
    """Load (once) and return the library; raises NativeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C binary-networks-pytorch_b200/csrc`). There is no non-CUDA fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def available() -> bool:
    return os.path.exists(LIB_PATH)


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().bnn_strerror(code)
        raise NativeError(f"{what} failed with code {code}: {msg.decode() if msg else '?'}")


def query(what: int) -> int:
    v = c_int64(0)
    check(lib().bnn_query(what, ctypes.byref(v)), "bnn_query")
    return int(v.value)


PLAN_FIELDS = ("P", "C", "kw_inst", "stride_inst", "csa", "TH", "TW", "warps", "units", "channel_tiles", "smem", "groups")


def conv_plan(geom: ConvGeom, flags: int = 0, sms: int = 0) -> dict:
    """Tile plan for a geometry (works without a GPU)."""
    arr = (c_int32 * 12)()
    check(lib().bnn_conv_plan(ctypes.byref(geom), flags, sms, arr), "bnn_conv_plan")
    return dict(zip(PLAN_FIELDS, list(arr)))


def conv_plan_list(geom: ConvGeom, flags: int = 0, cap: int = 4096) -> list:
    """Every feasible tile plan of a geometry, cost-model order (works without a GPU)."""
    arr = (c_int32 * (12 * cap))()
    n = c_int32(0)
    check(lib().bnn_conv_plan_list(ctypes.byref(geom), flags, arr, cap, ctypes.byref(n)), "bnn_conv_plan_list")
    return [dict(zip(PLAN_FIELDS, arr[12 * i: 12 * i + 12])) for i in range(min(cap, n.value))]


def conv_instance(geom: ConvGeom, ep: "Epilogue", flags: int, P: int, C: int, TH: int = 0, warps: int = 0) -> dict:
    """Kernel instance a forced-plan launch would run: P, C, kw / stride instance, carry-save mode, epilogue kind."""
    arr = (c_int32 * 6)()
    check(lib().bnn_conv_instance(ctypes.byref(geom), ctypes.byref(ep), flags, P, C, TH, warps, arr), "bnn_conv_instance")
    return dict(zip(("P", "C", "kw_inst", "stride_inst", "csa", "epi"), list(arr)))


def conv_split(geom: ConvGeom, flags: int = 0):
    """(chunks per part, number of parts) of the split-K decomposition; (all chunks, 1) when none is needed."""
    cpp, parts = c_int32(0), c_int32(0)
    check(lib().bnn_conv_split(ctypes.byref(geom), flags, ctypes.byref(cpp), ctypes.byref(parts)), "bnn_conv_split")
    return cpp.value, parts.value


def launch_count() -> int:
    return query(Q_LAUNCH_COUNT)
