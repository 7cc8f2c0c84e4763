"""In-tree build of ``libsmmregrid_b200.so`` (nvcc, sm_100a only).

The shared library is a plain C-ABI object (``include/smmregrid_b200.h``): no torch, no
pybind.  It is built into ``smmregrid_b200/lib/`` so it travels with the source tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
LIB_DIR = os.path.join(_PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsmmregrid_b200.so")
SOURCES = ["smm_api.cu", "smm_plan.cpp"]
HEADERS = ["smm_common.h", "smm_kernels.cuh", "smm_plan.h",
           os.path.join("..", "..", "include", "smmregrid_b200.h")]

NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: smmregrid_b200 needs the CUDA toolkit to build "
                       "(there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if sources are newer than the binary.  Returns its path."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout)
    if verbose:
        print(proc.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
