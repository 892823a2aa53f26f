#!/usr/bin/env python3
"""Build oracle/_ref/libref_sm100{,_fast}.so - the REFERENCE's own CUDA kernel for the B200.

MEASUREMENT INFRASTRUCTURE ONLY ("reference kernel on B200").  Uses the patched copy that
make_ref.py writes to oracle/_ref/kernel_patched.inc (signature-only patch), compiled by nvcc
against the real CUDA runtime.  Two variants: default flags, and --use_fast_math (the flag the
reference's project file sets, `Ray Tracer engine.vcxproj:66-69`).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref  # noqa: E402

ORACLE = os.path.dirname(HERE)
OUT = os.path.join(ORACLE, "_ref")
REF = make_ref.REF


def build(verbose=True):
    if not os.path.isfile(os.path.join(OUT, "kernel_patched.inc")):
        make_ref.build(verbose=False)
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    outs = []
    for name, extra in (("libref_sm100.so", []), ("libref_sm100_fast.so", ["--use_fast_math"])):
        so = os.path.join(OUT, name)
        cmd = ["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100,code=sm_100", "-rdc=true", "-w",
               "-Xcompiler", "-fPIC", "-shared", *extra,
               "-I", os.path.join(HERE, "stubs_gpu"), "-I", HERE, "-I", ORACLE, "-I", OUT, "-I", REF,
               os.path.join(HERE, "ref_gpu_driver.cu"), os.path.join(HERE, "sprite_raw.cpp"),
               os.path.join(REF, "memManager.cpp"), "-o", so]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd, env=env)
        outs.append(so)
    return outs


if __name__ == "__main__":
    print(build())
