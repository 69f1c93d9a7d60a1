"""TEST INFRASTRUCTURE — builds oracle/_build/libfn3_oracle.so from oracle/fn3_oracle.c (gcc).

The reference (davidhwyllie/findNeighbour3) is pure Python: there is no C/C++ source to
compile into oracle/_ref/, so no oracle/_ref is ever produced."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fn3_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libfn3_oracle.so")


def build_oracle(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-std=c11", "-Wall", "-shared", "-fPIC", "-pthread", "-o", LIB, SRC]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build_oracle(force=True))
