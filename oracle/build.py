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
