"""Build libgsage_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

    python graphsage-pytorch_b200/build.py [--force] [--verbose]

The library is plain CUDA behind a C ABI (include/gsage_b200.h); it cross-compiles on a
box without a GPU in a few seconds and travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgsage_b200.so")
STAMP = LIB + ".stamp"

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false" if False else "-Xptxas", "-O3",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the gsage_b200 CUDA library cannot be built")


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest() -> str:
    h = hashlib.sha256()
    files = _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
        [os.path.join(os.path.dirname(HERE), "include", "gsage_b200.h"), os.path.abspath(__file__)]
    for f in files:
        h.update(os.path.basename(f).encode())      # NOT the absolute path: the snapshot on the GPU box lives elsewhere
        with open(f, "rb") as fp:
            h.update(fp.read())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fp:
        return fp.read().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and is_current():
        return LIB
    import fcntl
    with open(LIB + ".lock", "w") as lock:          # ranks of one job may all find the library stale: one builds
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and is_current():
            return LIB
        tmp = f"{LIB}.{os.getpid()}.tmp"
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + _sources()
        if verbose:
            print(" ".join(cmd))
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
        if proc.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError("nvcc failed building libgsage_b200.so")
        os.replace(tmp, LIB)                        # atomic: a concurrent dlopen never sees a partial file
        with open(STAMP, "w") as fp:
            fp.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
