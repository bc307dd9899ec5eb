"""Build libdepthhead_cuda.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

    python -m depthhead_b200._build [--force]

Flags that matter for parity: -fmad=false (no FMA contraction on the device; Rust never fuses)
and -ffp-contract=off for the host parts (K^-1, kernel table).  -lineinfo so ncu's source page
maps to these files."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdepthhead_cuda.so")
SOURCES = ["dh_kernels.cu", "dh_ctx.cu", "dh_capi.cu", "dh_forest.cpp", "dh_biwi.cpp", "dh_train.cu", "dh_hostenc.cpp"]
HEADERS = ["dh_kernels.cuh", "dh_ctx.hpp", "dh_forest.hpp", "dh_json.hpp", "dh_types.hpp", "dh_hostenc.hpp",
           os.path.join("..", "..", "include", "depthhead_cuda.h")]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_id() -> str:
    """sha256 over the library's sources and headers, first 12 hex digits (dh_build_id())."""
    import hashlib
    hsh = hashlib.sha256()
    for f in sorted(SOURCES + HEADERS):
        with open(os.path.join(CSRC, f), "rb") as fh:
            hsh.update(f.encode() + b"\0" + fh.read() + b"\0")
    return hsh.hexdigest()[:12]


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2",
           "-shared", "-cudart", "static", "-Xcompiler", "-pthread", "-DDH_BUILD_ID=\"%s\"" % build_id(), "-o", LIB]
    if host_cxx:
        cmd += ["-ccbin", host_cxx]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, f) for f in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libdepthhead_cuda.so")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
