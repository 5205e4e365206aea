"""In-tree build of the sm_100a shared library (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc", "piso_b200.cu")
LIB_DIR = os.path.join(PKG, "_lib")
LIB = os.path.join(LIB_DIR, "libfluidgym_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-I" + os.path.join(REPO, "include")]


def lib_is_fresh() -> bool:
    if not os.path.exists(LIB):
        return False
    deps = [SRC, os.path.join(PKG, "csrc", "ortho3_b200.cuh"), os.path.join(PKG, "csrc", "extruded3_b200.cuh"),
            os.path.join(REPO, "include", "fluidgym_b200.h")]
    return all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and lib_is_fresh():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
