"""In-tree build of libore_b200.so (CUDA kernels + C ABI) for sm_100a.

`nvcc` cross-compiles without a GPU, so this runs in the build container; the resulting
.so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libore_b200.so")
SOURCES = ["ore_capi.cu", "ore_fast.cu"]
HEADERS = ["ore_kernels.cuh", "ore_device.cuh", "ore_libm.cuh", "ore_clusters.h", os.path.join("..", "..", "include", "ore_render.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # reference-exact sequences rely on unfused mul/add; filters use explicit fmaf
    "--fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise FileNotFoundError("nvcc not found; the render path has no non-CUDA fallback")


def is_stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build_library(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    if not force and not is_stale():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, *extra_flags, *[os.path.join(CSRC, s) for s in SOURCES], "-o", LIB_PATH]
    env = dict(os.environ)
    env.pop("CC", None)   # the image's CC/CXX point at a compiler without libgomp; nvcc uses the system g++
    env.pop("CXX", None)
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd, env=env)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
