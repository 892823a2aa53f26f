"""The C ABI header is valid C (not just C++), and the C++ drop-in shims + headless harness link against the
library (no GPU needed for either)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "use_abi.c"
    src.write_text(
        '#include "ore_render.h"\n'
        "int use(void) {\n"
        "    ore_context* ctx = 0;\n"
        "    ore_camera cam = {{4, 3, 10}, {0, 0, 1}, 0.f, 180.f, -20.f};\n"
        "    ore_frame fr = {640, 480, 0, 480, 1, 1.0f, ORE_FLAG_NONE, 0, 0};\n"
        "    unsigned int px[4];\n"
        "    if (ore_create(&ctx, 0) != ORE_OK) return 1;\n"
        "    return ore_render(ctx, &cam, &fr, px) + ore_destroy(ctx) + (int)sizeof(ore_counters);\n"
        "}\n")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "use_abi.o")])


def test_camera_struct_has_the_reference_layout(tmp_path):
    """36 bytes: Org@0, Dir@12, aspect@24, Camyaw@28, Campitch@32 (the by-value `camera` argument, kernel.cu:237-262)"""
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stddef.h>\n#include "ore_render.h"\n'
        "_Static_assert(sizeof(ore_camera) == 36, \"size\");\n"
        "_Static_assert(offsetof(ore_camera, dir) == 12 && offsetof(ore_camera, aspect) == 24, \"dir/aspect\");\n"
        "_Static_assert(offsetof(ore_camera, yaw) == 28 && offsetof(ore_camera, pitch) == 32, \"yaw/pitch\");\n"
        "int main(void) { return 0; }\n")
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "layout.o")])


def test_host_shims_and_headless_harness_link(pkg):
    pkg.build.build_library()
    host = os.path.join(ROOT, "ray-tracer-engine_b200", "host")
    subprocess.check_call(["make", "-s", "-C", host])
    assert os.path.isfile(os.path.join(host, "libore_host.so")) and os.path.isfile(os.path.join(host, "ore_headless"))
    syms = subprocess.run(["nm", "-D", "--defined-only", os.path.join(host, "libore_host.so")], capture_output=True, text=True).stdout
    # the reference's entry points with C++ linkage: void onStart(), void update() (kernel.cuh:3-4)
    assert "_Z7onStartv" in syms and "_Z6updatev" in syms
    # memManager::operator new/delete and check_cuda (memManager.h:11-18)
    assert "_ZN10memManagernwEm" in syms and "_ZN10memManagerdlEPv" in syms and "check_cuda" in syms
