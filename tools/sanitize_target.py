"""Small frames through every kernel generation, for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rte_b200  # noqa: E402

pkg = rte_b200.pkg
F = pkg.capi
r = pkg.Renderer(0)
for n, seed, W, H in ((64, 2, 161, 91), (1024, 3, 96, 54), (7000, 7, 48, 27), (0, 1, 40, 30)):
    sc = pkg.scene.scaled_scene(n, seed) if n else pkg.scene.reference_scene(0, 1)
    cam = pkg.scene.orbit_camera(sc, 11) if n else pkg.scene.reference_camera()
    r.set_scene(sc)
    for flags in (0, F.ORE_FLAG_FUSED_SHADOW, F.ORE_FLAG_EXHAUSTIVE, F.ORE_FLAG_COUNT_REFERENCE_TESTS):
        px = r.render(cam, W, H, flags=flags)
        r.render(cam, W, H, y0=3, y1=H, y_step=4, flags=flags)
    print(n, W, H, int(px.sum()), flush=True)
r.close()
print("done")
