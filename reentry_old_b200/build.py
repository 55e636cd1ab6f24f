"""In-tree build of the CUDA library (sm_100a only) -- `python -m reentry_old_b200.build`.

nvcc cross-compiles without a GPU; the resulting libreentry_b200.so stays in-tree (git-ignored,
but shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libreentry_b200.so")
SOURCES = [os.path.join(CSRC, "rb_api.cu")]
DEPS = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) + \
    [os.path.join(ROOT, "include", "reentry_b200.h"), os.path.join(ROOT, "include", "reentry_constants.h"),
     os.path.join(ROOT, "include", "rb_dop853_tables.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def command(extra=(), out=LIB):
    return [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
            "-fmad=false",  # every fused multiply-add is written out as fma(): identical arithmetic in every kernel instantiation
            "-Xcompiler", "-fPIC,-fopenmp,-O2", "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
            *extra, "-o", out, *SOURCES]


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in DEPS)


def build(force: bool = False, verbose: bool = False, extra=()) -> str:
    if force or is_stale():
        cmd = command(extra=(["-Xptxas", "-v"] if verbose else []) + list(extra))
        env = dict(os.environ)
        env.pop("CC", None)   # the image exports CC=/opt/gcc/bin/gcc, which has no libgomp
        env.pop("CXX", None)
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed building libreentry_b200.so")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
