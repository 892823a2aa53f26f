"""One-off sweep: many random awkward scenes, default kernels vs exhaustive mode (see tests/test_random_scenes_gpu.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rte_b200  # noqa: E402
from test_random_scenes_gpu import random_scene  # noqa: E402

pkg = rte_b200.pkg
r = pkg.Renderer(0)
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(lo, hi):
    sc, cam = random_scene(pkg, seed)
    r.set_scene(sc)
    W, H = 80 + seed % 37, 45 + seed % 11
    a = r.render(cam, W, H)
    ia, ta = r.hits(H, W)
    b = r.render(cam, W, H, flags=pkg.capi.ORE_FLAG_EXHAUSTIVE)
    ib, tb = r.hits(H, W)
    if not (np.array_equal(a, b) and np.array_equal(ia, ib) and np.array_equal(ta.view(np.uint32), tb.view(np.uint32))):
        bad += 1
        print("MISMATCH seed", seed, sc.name, int(np.count_nonzero(a != b)), "pixels", int(np.count_nonzero(ia != ib)), "ids", flush=True)
print(f"seeds {lo}..{hi - 1}: {bad} mismatching scenes")
