"""Quick check: ORE_FLAG_FAST_LIBM frames with the hit-count hint active (chunk pairs + catch-all launch) equal the
fused fast frame, at a size with several staging chunks."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rte_b200  # noqa: E402

t0 = time.time()
pkg = rte_b200.pkg
F = pkg.capi
sc = pkg.scene.scaled_scene(1024, 3)
cam = pkg.scene.orbit_camera(sc, 0, 240)
r = pkg.Renderer(0)
r.set_scene(sc)
W, H = 3840, 2160
a0 = r.render(cam, W, H, flags=F.ORE_FLAG_FAST_LIBM)
n0 = r.counters()["kernel_launches"]
a1 = r.render(cam, W, H, flags=F.ORE_FLAG_FAST_LIBM)
n1 = r.counters()["kernel_launches"]
b = r.render(cam, W, H, flags=F.ORE_FLAG_FAST_LIBM | F.ORE_FLAG_FUSED_SHADOW)
d = r.render(cam, W, H)
print("launches first/second frame", n0, n1, "fast==fast", np.array_equal(a0, a1), "fast==fused fast", np.array_equal(a1, b),
      "pixels differing from the bit-exact frame", int(np.count_nonzero(a1 != d)), "of", W * H, f"{time.time() - t0:.1f}s")
r.close()
