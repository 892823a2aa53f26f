#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the REFERENCE's own code (oracle/_ref/libref_oracle.so).

Run in the build container (where /root/reference exists):
    python oracle/ref_build/make_ref.py && python tests/golden/make_golden.py
The reference publishes no golden vectors, so these fixtures - pixels,
nearest-hit ids and the bit patterns of t for every case in tests/cases.py, plus single-ray
known answers - are what pins the C restatement (and, through it, the CUDA path) on
machines where the reference tree is absent.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, os.path.dirname(TESTS))
sys.path.insert(0, TESTS)
import cases  # noqa: E402
import oraclelib  # noqa: E402


def main():
    ref = oraclelib.load("ref")
    assert ref.kind == "reference"
    # triangle mesh fixture: a torus OBJ run through the reference's OWN loader + BVH builder
    import tempfile
    import rte_b200
    scn = rte_b200.pkg.scene
    with tempfile.TemporaryDirectory() as td:
        obj = os.path.join(td, "torus.obj")
        scn.write_torus_obj(obj)
        arr = ref.build_mesh(obj)
    np.savez_compressed(os.path.join(HERE, "mesh_torus.npz"), **arr)
    print("mesh_torus", arr["tris"].shape, "boxes", arr["box_bounds"].shape[0])
    for name, make, W, H, kw in cases.SMALL:
        sc, cam = make()
        out = ref.render(sc, cam, W, H, **kw)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            pixels=out["pixels"], ids=out["ids"], t_bits=out["t"].view(np.uint32))
        print(name, out["pixels"].shape, "hits", int((out["ids"] >= 0).sum()))
    # single-ray known answers through sphere::intersect
    rng = np.random.default_rng(12345)
    n = 4096
    org = rng.uniform(-12, 12, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    cen = rng.uniform(-10, 10, size=(n, 3)).astype(np.float32)
    # a third of the rays are aimed near the sphere so hits, grazes and inside-origins occur
    aim = rng.uniform(-0.3, 0.3, size=(n, 3)).astype(np.float32)
    k = n // 3
    v = (cen[:k] + aim[:k]) - org[:k]
    d[:k] = (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
    rad = (rng.uniform(0, 0.99, size=n).astype(np.float32)) ** 2
    org[-64:] = cen[-64:] + (aim[-64:] * np.float32(0.1))   # origins inside the sphere
    hit = np.zeros(n, dtype=np.uint8)
    tb = np.zeros(n, dtype=np.uint32)
    for i in range(n):
        h, t = ref.sphere_intersect(org[i], d[i], cen[i], float(rad[i]))
        hit[i] = h
        tb[i] = np.float32(t).view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "intersect_kat.npz"), org=org, dir=d, centre=cen, radius=rad, hit=hit, t_bits=tb)
    print("intersect_kat", n, "hits", int(hit.sum()))


if __name__ == "__main__":
    main()
