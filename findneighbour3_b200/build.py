"""Build recipe for the native pieces (run here on CPU: nvcc cross-compiles sm_100a).

    python -m findneighbour3_b200.build          # libfn3.so (CUDA engine + C-ABI)

The shared objects are built IN-TREE (findneighbour3_b200/libfn3.so) so that they
travel with a snapshot of the repository; nothing is JIT-compiled at import time.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libfn3.so")
SRC = [os.path.join(HERE, "csrc", "fn3_engine.cu")]
DEPS = sorted(glob.glob(os.path.join(HERE, "csrc", "*.h")) + glob.glob(os.path.join(HERE, "csrc", "*.cuh")))
HDR = [os.path.join(ROOT, "include", "fn3.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xptxas", "-v",
    "-shared", "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"),
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=True):
    """Compile libfn3.so for sm_100a.  Returns the path."""
    if not force and not _stale(LIB, SRC + HDR + DEPS):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libfn3.so")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + SRC
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "csrc", "build.log")
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + res.stdout)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("nvcc failed building libfn3.so (see %s)" % log)
    if verbose:
        for line in res.stdout.splitlines():
            if "registers" in line or "spill" in line and "0 bytes spill stores, 0 bytes spill loads" not in line:
                print("[nvcc]", line.strip())
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv))
