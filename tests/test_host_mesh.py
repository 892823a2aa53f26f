"""The shim's OBJ loader + flat-BVH builder (host/ore_mesh.cpp) against the reference's own (mesh::mesh and
createBvhMesh, kernel.cu:577-936, run through oracle/_ref): identical triangles, leaf boxes and leaf index lists."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "ray-tracer-engine_b200", "host")
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def host_lib(pkg):
    pkg.build.build_library()
    subprocess.check_call(["make", "-s", "-C", HOST])
    return C.CDLL(os.path.join(HOST, "libore_host.so"), mode=os.RTLD_LAZY)


def host_build(lib, path, cap_tris=1 << 16, cap_boxes=4096):
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    tris = np.zeros((cap_tris, 27), dtype=np.float32)
    bounds = np.zeros((cap_boxes, 6), dtype=np.float32)
    offs = np.zeros(cap_boxes + 1, dtype=np.int32)
    idx = np.zeros(cap_tris * 2, dtype=np.int32)
    nt, hn, nb = C.c_int(0), C.c_int(0), C.c_int(0)
    lib.ore_host_build_mesh.argtypes = [C.c_char_p, fp, C.c_int, ip, ip, fp, ip, C.c_int, ip, ip, C.c_int]
    rc = lib.ore_host_build_mesh(path.encode(), tris.ctypes.data_as(fp), cap_tris, C.byref(nt), C.byref(hn),
                                 bounds.ctypes.data_as(fp), offs.ctypes.data_as(ip), cap_boxes, C.byref(nb),
                                 idx.ctypes.data_as(ip), idx.size)
    assert rc == 0
    n, b = nt.value, nb.value
    return dict(tris=tris[:n].copy(), has_normals=bool(hn.value), box_bounds=bounds[:b].copy(),
                box_offsets=offs[: b + 1].copy(), box_indices=idx[: offs[b]].copy())


def same(a, b):
    return (a["has_normals"] == b["has_normals"] and np.array_equal(a["tris"].view(np.uint32), b["tris"].view(np.uint32))
            and np.array_equal(a["box_bounds"].view(np.uint32), b["box_bounds"].view(np.uint32))
            and np.array_equal(a["box_offsets"], b["box_offsets"]) and np.array_equal(a["box_indices"], b["box_indices"]))


def write_grid_obj(path, style, n=9, seed=3):
    """a bumpy n x n height field; style: 'vtn' a/b/c triangles, 'vtn_quads' a/b/c quads, 'vn' a//c triangles,
    'vn_quads' a//c quads, 'bare' plain index triangles"""
    rng = np.random.default_rng(seed)
    hgt = rng.uniform(0, 1.5, size=(n, n))
    with open(path, "w") as fh:
        for i in range(n):
            for j in range(n):
                fh.write("v %.5f %.5f %.5f\n" % (1 + 0.8 * i, 2 + hgt[i, j], 1 + 0.8 * j))
        if style.startswith("vtn"):
            for i in range(n):
                for j in range(n):
                    fh.write("vt %.5f %.5f\n" % (i / (n - 1), j / (n - 1)))
        if style != "bare":
            for i in range(n):
                for j in range(n):
                    a = math.atan2(hgt[i, j] - 0.7, 1.0)
                    fh.write("vn %.5f %.5f %.5f\n" % (math.sin(a) * 0.3, math.cos(a), math.sin(a) * 0.2))

        def corner(k):
            return {"vtn": "%d/%d/%d" % (k, k, k), "vtn_quads": "%d/%d/%d" % (k, k, k), "vn": "%d//%d" % (k, k),
                    "vn_quads": "%d//%d" % (k, k), "bare": "%d" % k}[style]

        for i in range(n - 1):
            for j in range(n - 1):
                a, b, c, d = i * n + j + 1, (i + 1) * n + j + 1, (i + 1) * n + j + 2, i * n + j + 2
                if style.endswith("quads"):
                    fh.write("f %s %s %s %s\n" % tuple(corner(k) for k in (a, b, c, d)))
                else:
                    fh.write("f %s %s %s\n" % tuple(corner(k) for k in (a, b, c)))
                    fh.write("f %s %s %s\n" % tuple(corner(k) for k in (a, c, d)))


def test_torus_matches_the_fixture_built_by_the_reference_loader(host_lib, pkg, tmp_path):
    obj = str(tmp_path / "torus.obj")
    pkg.scene.write_torus_obj(obj)
    got = host_build(host_lib, obj)
    want = dict(np.load(os.path.join(GOLDEN, "mesh_torus.npz")))
    want["has_normals"] = bool(want["has_normals"])
    assert same(got, want)


@pytest.mark.parametrize("style", ["vtn", "vtn_quads", "vn", "vn_quads", "bare"])
@pytest.mark.parametrize("n", [4, 9, 23])
def test_loader_and_bvh_match_the_reference(style, n, host_lib, oracle_ref, tmp_path):
    obj = str(tmp_path / f"grid_{style}_{n}.obj")
    write_grid_obj(obj, style, n=n)
    got = host_build(host_lib, obj)
    want = oracle_ref.build_mesh(obj, cap_tris=1 << 16)
    assert got["tris"].shape[0] > 0
    assert same(got, want), (style, n, got["tris"].shape, want["tris"].shape, got["box_bounds"].shape, want["box_bounds"].shape)
