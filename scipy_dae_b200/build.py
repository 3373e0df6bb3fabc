"""Build recipe for libb200dae.so (explicit nvcc for sm_100a, in-tree so the .so travels with
the repository snapshot to the GPU box)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libb200dae.so")
SOURCES = ["b200dae.cu"]
HEADERS = ["radau.cuh", "bdf.cuh", "fd_jacobian.cuh", "dense_lu.cuh", "fastmath.cuh", "problems.cuh",
           os.path.join("..", "..", "include", "b200dae.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libb200dae.so cannot be built")
    return nvcc


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return os.path.getmtime(LIB_PATH) < max(os.path.getmtime(d) for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/b200dae.cu -> scipy_dae_b200/libb200dae.so (cross-compiles without a GPU)."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
