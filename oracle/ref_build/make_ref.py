#!/usr/bin/env python3
"""Build oracle/_ref/libref_oracle.so - the REFERENCE's own per-pixel code as host C++.

TEST INFRASTRUCTURE ONLY.  Reads /root/reference/kernel.cu where it lies, writes a
mechanically patched copy to oracle/_ref/kernel_patched.inc (git-ignored; reference
sources are never committed) and compiles it with the stand-in headers in stubs/.

The patch does not touch arithmetic:
  * `vec3d&` -> `const vec3d&` in the 10 non-mutating vector helpers, camera::rotateDir
    and reflect (MSVC binds temporaries to non-const references; g++/nvcc do not);
  * a forwarding overload `normalise(vec3d&&)` (the temporary is mutated and returned,
    which is what MSVC does with the original);
  * update() (the only `<<< >>>` launch) is fenced by `#ifdef __CUDACC__`.
Every edit asserts the original text of the line it changes, so a different reference
revision fails loudly instead of being silently mis-patched.

Compiler flags: -O2 -ffp-contract=off (no FMA contraction, no -ffast-math), OpenMP.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
OUT = os.path.join(ORACLE, "_ref")
REF = os.environ.get("ORE_REFERENCE_DIR", "/root/reference")

# (1-based line, regex that must match the original line)
CONST_REF_LINES = {
    46: r"vec3d sub\(vec3d& vec1, vec3d& vec2\)",
    51: r"vec3d sub\(float k, vec3d& vec2\)",
    56: r"vec3d divide\(vec3d& vec1, float k\)",
    61: r"vec3d add\(vec3d& vec1, vec3d& vec2\)",
    66: r"vec3d add\(vec3d& vec1, float a\)",
    71: r"vec3d multiply\(vec3d& vec1, vec3d& vec2\)",
    81: r"vec3d cross\(vec3d& vec1, vec3d& vec2\)",
    88: r"float dotproduct\(vec3d& vec1\)",
    93: r"float dotproduct\(vec3d& vec1, vec3d& vec2\)",
    98: r"float length\(vec3d& vec\)",
    248: r"vec3d rotateDir\(vec3d &vec,float yaw,float pitch\)",
    1283: r"vec3d reflect\(vec3d &I, vec3d &N\)",
}
NORMALISE_END = (108, r"^\}\s*$")              # closing brace of normalise(vec3d&)
UPDATE_BEGIN = (1762, r"^void update\(\) \{")
UPDATE_END = (1792, r"^\}\s*$")


def patch(src_lines):
    out = []
    for i, line in enumerate(src_lines, start=1):
        if i in CONST_REF_LINES:
            if not re.search(CONST_REF_LINES[i], line):
                raise SystemExit(f"make_ref: kernel.cu:{i} is not the expected signature: {line!r}")
            line = re.sub(r"vec3d\s*&", "const vec3d& ", line)
        if i == UPDATE_BEGIN[0]:
            if not re.search(UPDATE_BEGIN[1], line):
                raise SystemExit(f"make_ref: kernel.cu:{i} is not the start of update(): {line!r}")
            out.append("#ifdef __CUDACC__\n")
        out.append(line)
        if i == NORMALISE_END[0]:
            if not re.search(NORMALISE_END[1], line):
                raise SystemExit(f"make_ref: kernel.cu:{i} is not the end of normalise(): {line!r}")
            out.append("__device__ __host__\nvec3d normalise(vec3d&& v) { return normalise(v); }\n")
        if i == UPDATE_END[0]:
            if not re.search(UPDATE_END[1], line):
                raise SystemExit(f"make_ref: kernel.cu:{i} is not the end of update(): {line!r}")
            out.append("#endif\n")
    return out


def build(verbose=True):
    src = os.path.join(REF, "kernel.cu")
    if not os.path.isfile(src):
        raise SystemExit(f"make_ref: {src} not found (the reference only exists in the build container)")
    os.makedirs(OUT, exist_ok=True)
    with open(src, encoding="utf-8", errors="replace") as fh:
        lines = fh.readlines()
    patched = os.path.join(OUT, "kernel_patched.inc")
    with open(patched, "w", encoding="utf-8") as fh:
        fh.writelines(patch(lines))
    so = os.path.join(OUT, "libref_oracle.so")
    cmd = [
        "g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-w",
        "-I", os.path.join(HERE, "stubs"),      # stand-ins win over system/CUDA headers
        "-I", HERE, "-I", ORACLE, "-I", OUT,
        "-I", REF,                              # sprite.h, memManager.h, window.h, kernel.cuh in place
        os.path.join(HERE, "ref_driver.cpp"),
        os.path.join(HERE, "sprite_raw.cpp"),
        os.path.join(REF, "memManager.cpp"),    # compiled unmodified against the stub runtime
        "-o", so,
    ]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    print(build())
