"""The C-ABI library loads and exports exactly what include/bnn_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

import bnn_b200
from bnn_b200 import native

HEADER = os.path.join(ROOT, "include", "bnn_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bnn_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_functions() == native.exported_symbols()


def test_library_exports_every_declared_symbol():
    assert native.available(), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = native.lib()
    for name in _declared_functions():
        assert hasattr(lib, name), name


def test_size_helpers_and_errors_without_gpu():
    lib = native.lib()
    assert lib.bnn_act_bits_bytes(2, 64, 3, 5) == 2 * 1 * 3 * 5 * 16
    assert lib.bnn_act_bits_bytes(1, 65, 1, 1) == 2 * 16
    assert lib.bnn_weight_bits_bytes(33, 64, 3, 3) == 2 * 9 * 32 * 8
    assert lib.bnn_act_bits_bytes(0, 1, 1, 1) == 0
    assert native.query(native.Q_ABI_VERSION) == 4
    assert native.query(native.Q_SM_ARCH) == 100
    assert b"NULL" in lib.bnn_strerror(-1)
    # argument errors are reported before any CUDA call
    assert lib.bnn_pack_act_f32(None, 0, 0, 0, 0, 1, 1, 1, 1, None, None, 0, None, None) == -1
    assert lib.bnn_pack_weight_f32(None, 1, 1, 1, 1, 0, 1, None, None, None, None) == -1
    assert lib.bnn_bconv2d_fwd(None, None, None, None, None, None, 0, 0, 0, 0, None, 0, None) == -1
    g = native.ConvGeom(1, 64, 4, 4, 64, 3, 3, 0, 1, 1, 1, 1, 1)   # stride 0
    one = ctypes.c_void_p(16)
    assert lib.bnn_bconv2d_fwd(one, one, None, None, None, one, 0, 0, 0, 0, ctypes.byref(g), 0, None) == -2
    with pytest.raises(native.NativeError):
        native.check(-3, "x")
    # the entry points added in ABI v3 validate the same way
    assert lib.bnn_shortcut_fwd(None, 0, 0, 0, 1, 64, 8, 8, 2, 1, None, 64, None, None, None, None, None, None, 0, None) == -1
    assert lib.bnn_shortcut_fwd(one, 0, 0, 0, 1, 64, 8, 8, 0, 1, one, 64, None, None, None, None, None, one, 0, None) == -2   # pool 0
    assert lib.bnn_shortcut_fwd(one, 0, 0, 0, 1, 64, 8, 8, 2, 1, one, 64, None, None, None, one, None, one, 0, None) == -1   # bn_scale without bn_shift
    assert lib.bnn_stem_mma_weight_bytes() == 11 * 8 * 32 * 16
    assert lib.bnn_stem_mma_pack_weight(None, 0, None, None) == -1
    assert lib.bnn_stem_mma_pack_weight(one, 99, one, None) == -2
    assert lib.bnn_stem_mma_fwd(None, 1, 8, 8, None, 7, None, 0, None, None, None, None, None, None, 0, None) == -1
    assert lib.bnn_stem_mma_fwd(one, 1, 5, 5, one, 7, None, 0, one, one, None, None, one, None, 0, None) == -2              # smaller than the kernel
    # ABI v4: tcgen05 stem, input-range guard, planner hooks
    assert lib.bnn_stem_tc_weight_bytes() == 12 * 4096
    assert lib.bnn_stem_tc_pack_weight(None, 0, None, None) == -1
    assert lib.bnn_stem_tc_pack_weight(one, 99, one, None) == -2
    assert lib.bnn_stem_tc_fwd(None, 1, 8, 8, None, 7, None, 0, None, None, None, None, None, None, 0, None) == -1
    assert lib.bnn_stem_tc_fwd(one, 1, 5, 5, one, 7, None, 0, one, one, None, None, one, None, 0, None) == -2
    assert lib.bnn_stem_tc_fwd(one, 1, 8, 8, one, 7, None, 0, one, one, one, None, one, None, 0, None) == -1               # nx_scale without nx_shift
    assert lib.bnn_amax_f32(None, 4, None, None) == -1
    assert lib.bnn_amax_f32(one, 0, one, None) == -2
    assert lib.bnn_amax_f32(ctypes.c_void_p(4), 4, one, None) == -5
    n = ctypes.c_int32(0)
    g3 = native.ConvGeom(2, 64, 8, 56, 128, 3, 3, 1, 1, 1, 1, 1, 1)
    assert lib.bnn_conv_plan_list(ctypes.byref(g3), 0, None, 0, ctypes.byref(n)) == 0 and n.value > 10
    assert lib.bnn_conv_plan_list(None, 0, None, 0, ctypes.byref(n)) == -1
    ep = native.Epilogue()
    assert lib.bnn_bconv2d_fused_fwd_plan(one, one, ctypes.byref(g3), ctypes.byref(ep), 0, 5, 3, 0, 0, None) == -3       # no such family
    # split-K decomposition (host only): the VGG classifier shape needs it, a ResNet layer does not
    assert native.conv_split(native.ConvGeom(1, 25088, 1, 8, 4096, 1, 1, 1, 1, 0, 0, 1, 1))[1] > 1
    assert native.conv_split(native.ConvGeom(256, 512, 7, 7, 512, 3, 3, 1, 1, 1, 1, 1, 1)) == (8, 1)
    cpp, parts = native.conv_split(native.ConvGeom(1, 16448, 4, 4, 64, 1, 1, 1, 1, 0, 0, 1, 1))
    assert parts > 1 and cpp * parts >= 257 and cpp * (parts - 1) < 257 and cpp <= 256
    assert lib.bnn_bconv2d_partial_fwd(one, one, ctypes.byref(g3), 0, 2, one, 0, None) == -2        # only one chunk exists
    assert lib.bnn_bconv2d_partial_fwd(one, one, ctypes.byref(g3), 0, 1, None, 0, None) == -1
    assert lib.bnn_dot_finish_f32(one, 2, 3, None, None, None, one, 0, 0, 0, 0, 1, 1, 1, 1, None) == -2   # divisor 1 or 2
    assert lib.bnn_dot_finish_f32(None, 1, 1, None, None, None, one, 0, 0, 0, 0, 1, 1, 1, 1, None) == -1
    assert lib.bnn_pack_weight_ternary_f32(None, 1, 1, 1, 1, 0, 1, None, None, None, None, None) == -1
    inst = (ctypes.c_int32 * 6)()
    assert lib.bnn_conv_instance(ctypes.byref(g3), ctypes.byref(ep), 0, 8, 2, 0, 0, inst) == 0 and list(inst) == [8, 2, 3, 1, 1, 0]
    assert lib.bnn_stem_fwd(None, 1, 8, 8, None, None, None, None, None, None, None, 0, None) == -1
    hp, wp = ctypes.c_int32(0), ctypes.c_int32(0)
    assert lib.bnn_stem_out_hw(224, 224, ctypes.byref(hp), ctypes.byref(wp)) == 0 and (hp.value, wp.value) == (56, 56)


def test_sass_contains_tma_and_popc():
    """Evidence the shipped kernels are the sm_100a TMA + POPC design (SASS mnemonics per
    B200_PROFILING.md); skipped when cuobjdump is unavailable."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump")
    sass = subprocess.run([cuobjdump, "-sass", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTMALDG", "UBLKCP", "POPC", "LOP3", "HMMA.16816.F32", "FFMA2",      # + the mma.sync / fma stems
                     "UTCHMMA", "LDTM", "UTCBAR",                                         # tcgen05 stem: MMA, TMEM load, commit
                     "ACQBULK", "PREEXIT"):                                               # programmatic dependent launch
        assert mnemonic in sass, mnemonic
    # the lean fused epilogue (EPI 3 instances) converts its integer dots without the conversion pipe, which POPC saturates
    blocks = sass.split("Function : ")[1:]
    lean = [b for b in blocks if b.startswith("_ZN3bnn12bconv_kernel") and b.split("\n", 1)[0].rstrip().endswith("ELi3EEEv14CUtensorMap_stNS_8ConvArgsE")]
    assert len(lean) >= 20
    assert not any("I2FP.F32.S32" in b for b in lean)
