"""Builds libgm2.so in-tree with nvcc for sm_100a (the only target)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
SRC = os.path.join(PKG_DIR, "csrc", "gm2.cu")
HOST_SRC = [os.path.join(PKG_DIR, "csrc", "host_expand.cpp")]      # host-only C++, compiled by nvcc's host compiler
HDR = os.path.join(REPO_ROOT, "include", "gm2.h")
LIB = os.environ.get("GM2_LIB") or os.path.join(PKG_DIR, "libgm2.so")     # GM2_LIB: another build of the library (A/B runs)

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    csrc = os.path.dirname(SRC)
    deps = [HDR] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".cpp", ".hpp"))]
    return any(os.path.getmtime(p) > t for p in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [find_nvcc(), *NVCC_FLAGS, "-I", os.path.join(REPO_ROOT, "include"), "-o", LIB, SRC, *HOST_SRC]
    if os.environ.get("GM2_EMIT_DEBUG"):          # timing knock-outs in k_emit (wrong output): experiments only
        cmd.insert(1, "-DGM2_EMIT_DEBUG")
    for flag in os.environ.get("GM2_NVCC_EXTRA", "").split():      # e.g. -DGM2_EXP_...=1 for an experiment build
        cmd.insert(1, flag)
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
