"""Builds the C-ABI shared library in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

The translation units (csrc/*.cu) are compiled in parallel and linked into ``libnind_b200.so``.  The library is
rebuilt when any source / header or the compile flags (incl. ``NIND_NVCC_EXTRA``) differ from what the existing
library was built from: a hash of all of them is stored next to it (``libnind_b200.so.key``).
"""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnind_b200.so")
KEY_PATH = LIB_PATH + ".key"
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))


def headers():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh"))) + [os.path.join(HERE, "..", "include", "nind_b200.h")]


def extra_flags():
    # NIND_NVCC_EXTRA="-DNIND_SETS_PM=4" etc.: development builds of kernel variants
    return os.environ.get("NIND_NVCC_EXTRA", "").split()


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libnind_b200.so")


def build_key() -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS + extra_flags()).encode())
    for p in sources() + headers():
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(KEY_PATH):
        return True
    try:
        return open(KEY_PATH).read().strip() != build_key()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> libnind_b200.so (sm_100a).  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    flags = NVCC_FLAGS + extra_flags()

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(src))[0] + ".o")
        cmd = [nvcc] + flags + ["-c", "-o", obj, src]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + r.stdout + r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources()))) as pool:
        objs = list(pool.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
    with open(KEY_PATH, "w") as f:
        f.write(build_key())
    return LIB_PATH
