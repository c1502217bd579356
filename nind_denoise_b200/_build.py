"""Builds the C-ABI shared library in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnind_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "net.cu")]
HEADERS = [os.path.join(HERE, "csrc", h) for h in ("ptx.cuh", "igemm.cuh", "igemm_host.cuh", "aux.cuh")] + [
    os.path.join(HERE, "..", "include", "nind_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libnind_b200.so")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS if os.path.exists(p))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> libnind_b200.so (sm_100a).  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    # NIND_NVCC_EXTRA="-DNIND_PAIR_MODE=1" compiles experimental code paths in (development only)
    cmd = [find_nvcc()] + NVCC_FLAGS + os.environ.get("NIND_NVCC_EXTRA", "").split() + ["-o", LIB_PATH] + SOURCES
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return LIB_PATH
