"""In-tree build of libplonkish_cuda.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import hashlib
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = [os.path.join(_HERE, "csrc", "api.cu"), os.path.join(_HERE, "csrc", "host_copy.cpp")]
HEADERS = [
    os.path.join(_HERE, "csrc", h) for h in ("fq.cuh", "g1.cuh", "msm_kernels.cuh", "poly_kernels.cuh", "sumcheck_kernels.cuh", "lookup_kernels.cuh", "dpfq.cuh")
] + [os.path.join(os.path.dirname(_HERE), "include", "plonkish_cuda.h")]
OUT = os.path.join(_HERE, "libplonkish_cuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


STAMP = OUT + ".stamp"


def _fingerprint() -> str:
    """sha256 over the compiler flags and every source / header: the library is rebuilt when its sources changed, however
    the files' mtimes came out of a checkout or a copy (an mtime gate can skip a needed build)."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in SOURCES + HEADERS:
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_stale() -> bool:
    if not os.path.exists(OUT) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _fingerprint()


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    stamp = _fingerprint()  # of what the compiler is about to read: a source edited during the build leaves the library stale
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    with open(STAMP, "w") as f:
        f.write(stamp + "\n")
    if verbose:
        print(res.stderr)
    return OUT
