"""Build libccb200.so (hand-written sm_100a CUDA + the C ABI of include/ccb200.h) in-tree.

    python -m coveragecontrol_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libccb200.so")
SOURCES = ["ccb200.cu"]
HEADERS = ["common.cuh", "idf.cuh", "voronoi.cuh", "vglobal.cuh", "vwindow.cuh", "maps.cuh", "obs.cuh"]

NVCC_FLAGS = [
    "-std=c++17",
    "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",          # explicit fma()/fmaf() only: see ccb200.cu header comment
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: coveragecontrol_b200 needs the CUDA toolkit to build")


HASH = LIB + ".srchash"


def _source_hash() -> str:
    """sha256 over every file the library is built from and the flags: staleness by CONTENT, not by mtime (a copied
    tree -- the GPU box gets a snapshot -- need not keep the order of modification times)."""
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps += [os.path.join(_HERE, "..", "include", "ccb200.h"), os.path.join(_HERE, "..", "include", "ccb200_test.h")]
    for d in deps:
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(HASH):
        return True
    with open(HASH) as f:
        return f.read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    import fcntl

    # one builder at a time (every rank of a torchrun may get here at once); the others find the library fresh
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB
            tmp = f"{LIB}.{os.getpid()}.tmp"
            cmd = [_nvcc(), *NVCC_FLAGS, "-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
            out = subprocess.run(cmd, capture_output=True, text=True)
            log = out.stdout + out.stderr
            with open(os.path.join(_HERE, "build.log"), "w") as f:
                f.write(" ".join(cmd) + "\n" + log)
            if out.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + log)
            os.replace(tmp, LIB)  # atomic: a concurrent dlopen sees the old or the new file, never a partial one
            with open(HASH, "w") as f:
                f.write(_source_hash() + "\n")
            if verbose:
                print(log)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
