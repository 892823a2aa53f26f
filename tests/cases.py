"""Shared parity cases: (name, scene factory, camera factory, W, H, render kwargs)."""
import rte_b200

scene = rte_b200.pkg.scene


def _scaled(n, seed, frame):
    sc = scene.scaled_scene(n, seed)
    return sc, scene.orbit_camera(sc, frame)


def _ref(n, seed, checker=False):
    if checker:
        sc = scene.reference_scene(n, seed, texture=scene.checker_texture(256, 256), sky=scene.checker_texture(512, 256, 32))
    else:
        sc = scene.reference_scene(n, seed)
    return sc, scene.reference_camera()


def _prims(n, seed, frame, n_cubes, cseed, plane=True):
    base = scene.scaled_scene(n, seed) if n else scene.reference_scene(0, 1)
    sc = scene.with_cubes_and_plane(base, n_cubes, cseed, plane=plane)
    return sc, scene.orbit_camera(sc, frame)


def torus_mesh():
    import os
    import numpy as np
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mesh_torus.npz"))
    return scene.Mesh.from_arrays(d)


def _mesh(n, seed, frame, cubes=0):
    sc = scene.reference_scene(n, seed)
    if cubes:
        sc = scene.with_cubes_and_plane(sc, cubes, 5)
    sc.mesh = torus_mesh()
    cam = scene.reference_camera() if frame is None else scene.orbit_camera(sc, frame)
    return sc, cam


def _refprims():
    sc = scene.with_cubes_and_plane(scene.reference_scene(64, 1), 8, 1, reference_formula=True)
    return sc, scene.reference_camera()


# small cases: the oracle finishes each in well under a second
SMALL = [
    ("R64_refcam_160x120", lambda: _ref(64, 1), 160, 120, {}),
    ("R64_checker_128x96", lambda: _ref(64, 1, checker=True), 128, 96, {}),
    ("S64_f0_161x91", lambda: _scaled(64, 2, 0), 161, 91, {}),            # ragged width (not a multiple of 32)
    ("S64_f77_96x54", lambda: _scaled(64, 2, 77), 96, 54, {}),
    ("R1024_inside_96x54", lambda: _ref(1024, 1), 96, 54, {}),              # camera inside the sphere cloud: negative-t hits
    ("S1024_f0_128x72", lambda: _scaled(1024, 3, 0), 128, 72, {}),
    ("S1024_f100_band", lambda: _scaled(1024, 3, 100), 3840, 2160, {"y0": 700, "y1": 1500, "y_step": 199}),
    ("S7000_f30_48x27", lambda: _scaled(7000, 7, 30), 48, 27, {}),         # streaming tiles, ragged last chunk
    ("S16384_f0_64x36", lambda: _scaled(16384, 5, 0), 64, 36, {}),         # config-5 scene, >1 smem tile pass
    ("R1_one_sphere_64x48", lambda: _ref(1, 9), 64, 48, {}),
    ("R0_empty_scene_64x48", lambda: _ref(0, 1), 64, 48, {}),              # empty input: sky only
    ("R3_single_row", lambda: _ref(3, 4), 257, 480, {"y0": 240, "y1": 241}),
    # cubes and the plane: hit types 3 and 2, shadows from all three primitive kinds
    ("S64_cubes_plane_f0_160x90", lambda: _prims(64, 2, 0, 12, 5), 160, 90, {}),
    ("S64_cubes_plane_f120_96x54", lambda: _prims(64, 2, 120, 12, 5), 96, 54, {}),
    ("R64_refcubes_plane_128x96", lambda: _refprims(), 128, 96, {}),
    ("cubes_only_plane_96x72", lambda: _prims(0, 1, 10, 5, 9), 96, 72, {}),
    # triangle mesh: a torus loaded by the reference's own OBJ loader / BVH builder (fixture)
    ("R64_torus_refcam_160x120", lambda: _mesh(64, 1, None), 160, 120, {}),
    ("R64_torus_orbit70_128x96", lambda: _mesh(64, 1, 70), 128, 96, {}),
    ("R16_cubes_plane_torus_128x96", lambda: _mesh(16, 3, 30, cubes=4), 128, 96, {}),
    ("torus_only_96x72", lambda: _mesh(0, 1, 100), 96, 72, {}),
]


def fnv1a(arr) -> int:
    h = 0xCBF29CE484222325
    data = arr.tobytes()
    # vectorised FNV is awkward; frames here are small enough for the plain loop in C via hashlib-free path
    for b in data:
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h
