"""Build libequilibrium_cuda.so in-tree with nvcc for sm_100a.

    python -m equilibrium_b200.build [--force]

nvcc cross-compiles without a GPU.  -fmad=false: the exact mode must round every
multiply and add separately like the reference (the kernels also use explicit
__fadd_rn/__fmul_rn, the flag covers everything else).
"""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "libequilibrium_cuda.so")
NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    hdr = os.path.join(_HERE, "..", "include", "equilibrium_cuda.h")
    return any(os.path.getmtime(s) > t for s in sources() + [hdr])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", OUT, os.path.join(CSRC, "eq_api.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
