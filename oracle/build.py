"""Compile oracle/bnn_oracle.c into oracle/liboracle_bnn.so with plain gcc (TEST INFRASTRUCTURE)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "bnn_oracle.c")
LIB = os.path.join(HERE, "liboracle_bnn.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fno-fast-math",
                               "-ffp-contract=off", "-o", LIB, SRC, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force=True))


# ----------------------------------------------------------------------------------------------------------------
# oracle/_ref: the UNMODIFIED reference, byte-compiled from the sources where they lie under /root/reference.
# The reference is pure Python (no native code to compile), so "building" it is `py_compile`: every module of its
# `bnn` package becomes a sourceless .pyc under oracle/_ref/bnn_ref/ (same interpreter on the GPU box: same image).
# No reference SOURCE is copied into the repo; oracle/_ref/ is git-ignored and travels with the snapshot like the
# built .so files.  Users: tests (validating the restatement), bench.py --impl reference (kind "reference").
# ----------------------------------------------------------------------------------------------------------------
REF_SRC = "/root/reference/bnn"
REF_OUT = os.path.join(HERE, "_ref")
REF_PKG = os.path.join(REF_OUT, "bnn_ref")


def build_ref(force: bool = False) -> str:
    """Byte-compile /root/reference/bnn -> oracle/_ref/bnn_ref/**/*.pyc.  No-op (returns the existing directory or "")
    when the reference tree is not present (the GPU box uses the prebuilt files)."""
    import py_compile
    if not os.path.isdir(REF_SRC):
        return REF_PKG if os.path.isdir(REF_PKG) else ""
    for root, _dirs, files in os.walk(REF_SRC):
        rel = os.path.relpath(root, REF_SRC)
        out_dir = os.path.normpath(os.path.join(REF_PKG, rel))
        for f in files:
            if not f.endswith(".py"):
                continue
            src, dst = os.path.join(root, f), os.path.join(out_dir, f + "c")
            if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
                os.makedirs(out_dir, exist_ok=True)
                py_compile.compile(src, cfile=dst, dfile=os.path.join("bnn", rel, f), doraise=True,
                                   invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    return REF_PKG


def load_ref():
    """Import the byte-compiled reference as package ``bnn_ref`` (None if oracle/_ref has not been built)."""
    import importlib.util
    import sys
    if "bnn_ref" in sys.modules:
        return sys.modules["bnn_ref"]
    init = os.path.join(REF_PKG, "__init__.pyc")
    if not os.path.exists(init):
        return None
    spec = importlib.util.spec_from_file_location("bnn_ref", init, submodule_search_locations=[REF_PKG])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["bnn_ref"] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        del sys.modules["bnn_ref"]
        raise
    return mod
