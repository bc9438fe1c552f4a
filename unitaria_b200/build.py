"""Builds ``libunitaria_b200.so`` (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

The library has no Python or torch dependency; it is loaded through ctypes (unitaria_b200/cabi.py).
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libunitaria_b200.so")
SOURCES = ["b2sv.cu"]
HEADERS = ["kernels.cuh", "planner.hpp", "tileops.hpp", "perm_jit.hpp", os.path.join("..", "..", "include", "b2sv.h")]

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-std=c++17",
    "-shared",
    "-Xcompiler",
    "-fPIC,-O3,-Wall,-Wno-unknown-pragmas",
    "--expt-relaxed-constexpr",
    "-ldl",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libunitaria_b200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    lib_mtime = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > lib_mtime for d in deps if os.path.exists(d))


DBG_LIB_PATH = os.path.join(_HERE, "libunitaria_b200_dbg.so")


def build(force: bool = False, verbose: bool = False, bounds: bool = False) -> str:
    """``bounds``: the range-checking debug build (-DB2SV_BOUNDS_CHECK -> libunitaria_b200_dbg.so; load it with
    B2SV_LIB=<path>), the stand-in for compute-sanitizer's memcheck on pools where the tool is closed."""
    if bounds:
        cmd = [find_nvcc()] + NVCC_FLAGS + ["-DB2SV_BOUNDS_CHECK", "-o", DBG_LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
            raise RuntimeError("nvcc failed building libunitaria_b200_dbg.so")
        return DBG_LIB_PATH
    if not force and not needs_build():
        return LIB_PATH
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])
    cmd += ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libunitaria_b200.so")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, bounds="--bounds" in sys.argv))
