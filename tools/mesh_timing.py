"""Timing of the 'next' primitives on one GPU: spheres + cubes + plane + triangle mesh (torus fixture)."""
import os
import sys
import statistics

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import rte_b200  # noqa: E402

pkg = rte_b200.pkg
scene = pkg.scene
r = pkg.Renderer(0)
for name, sc in (("R(64,1)", scene.reference_scene(64, 1)),
                 ("R(64,1)+12 cubes+plane", scene.with_cubes_and_plane(scene.reference_scene(64, 1), 12, 5)),
                 ("R(64,1)+torus(576 tris,127 leaves)", None), ("R(64,1)+cubes+plane+torus", None)):
    if sc is None:
        sc = scene.reference_scene(64, 1) if "cubes" not in name else scene.with_cubes_and_plane(scene.reference_scene(64, 1), 12, 5)
        sc.mesh = cases.torus_mesh()
    r.set_scene(sc)
    for W, H in ((1920, 1080), (3840, 2160)):
        ms = []
        for f in range(5):
            r.render(scene.orbit_camera(sc, f * 20), W, H)
            if f >= 1:
                ms.append(r.kernel_ms()[:3])
        c = r.counters()
        tot = statistics.mean(sum(m) for m in ms)
        print(f"{name:36s} {W}x{H}: primary {statistics.mean(m[1] for m in ms):.3f} ms shadow {statistics.mean(m[2] for m in ms):.3f} ms "
              f"-> {W * H / tot / 1e3:.0f} Mrays/s, hit {c['hit_pixels'] / c['pixels']:.2f}", flush=True)
r.close()
