"""Builds csrc/libcannoles_b200.so with nvcc for sm_100a (in-tree, so it travels with gpurun)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libcannoles_b200.so")
SOURCES = ["engine.cu", "capi.cu", "batched.cu", "measure.cu", "symbolic.cpp", "ordering.cpp"]
HEADERS = ["kernels.cuh", "batched_kernels.cuh", "nls_kernels.cuh", "ldlt_packed.cuh", "mma.cuh", "engine.h", "plan.h", "symbolic.h", "b2_cuda.h",
           os.path.join("..", "..", "include", "cannoles_b200.h")]
METIS = "/usr/local/cuda/lib64/libmetis_static.a"


def nvcc_path() -> str:
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
           "-std=c++17", "-Xcompiler", "-fPIC,-O3", "-shared", "-o", OUT]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, f) for f in SOURCES]
    cmd += [METIS, "-lcublas", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc build of libcannoles_b200.so failed")
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
