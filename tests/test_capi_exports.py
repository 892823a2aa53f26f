"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/ore_render.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "ore_render.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ore_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface(pkg):
    assert sorted(pkg.capi.EXPORTS) == declared_functions()


def test_library_builds_and_exports_all_symbols(pkg):
    path = pkg.build.build_library()
    lib = ctypes.CDLL(path)
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.ore_abi_version() == 2


def test_library_carries_only_an_sm100a_image(pkg):
    import subprocess
    path = pkg.build.build_library()
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out)


def test_no_gpu_means_loud_failure_not_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.OreError):
        pkg.Renderer(0)


def test_product_code_never_touches_the_oracle():
    """Nothing under the package or include/ may reference oracle/ (it is test infrastructure)."""
    pk = os.path.join(ROOT, "ray-tracer-engine_b200")
    for base, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "oracle/" not in txt and "liboracle" not in txt and "oraclelib" not in txt, os.path.join(base, f)


def test_python_flag_constants_match_the_header(pkg):
    text = open(os.path.join(ROOT, "include", "ore_render.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    header = {k: int(v) for k, v in re.findall(r"\b(ORE_FLAG_[A-Z_]+)\s*=\s*(\d+)", text)}
    assert len(header) >= 6
    for name, value in header.items():
        if name == "ORE_FLAG_NONE":
            continue
        assert getattr(pkg.capi, name) == value, name
    values = [v for k, v in header.items() if k != "ORE_FLAG_NONE"]
    assert len(set(values)) == len(values) and all(v & (v - 1) == 0 for v in values), "flags must be distinct bits"
