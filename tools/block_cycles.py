"""Per-block SM clocks of the two shadow-pass kernels for one frame (ORE_DEBUG_BLOCK_CYCLES hook): distribution + tail."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
path = "/tmp/ore_block_cycles.bin"
os.environ["ORE_DEBUG_BLOCK_CYCLES"] = path
import bench  # noqa: E402
import rte_b200  # noqa: E402

pkg = rte_b200.pkg
wl = sys.argv[1] if len(sys.argv) > 1 else "8k1024"
W, H, sc, camera, desc = bench.make_workload(pkg, wl)
r = pkg.Renderer(0)
r.set_scene(sc)
for f in range(3):
    r.render(camera(f), W, H)
hits = r.counters()["hit_pixels"]
r.close()
raw = np.fromfile(path, dtype=np.uint8)
cap = int(np.frombuffer(raw[:8], dtype=np.uint64)[0])
d = np.frombuffer(raw[8:], dtype=np.uint32).reshape(2, cap)
nb = (hits + 31) // 32
for name, v in (("stage A", d[0, :nb]), ("stage B", d[1, :nb])):
    v = v.astype(np.float64) / 1.965e3   # microseconds at 1965 MHz
    q = np.percentile(v, [1, 10, 50, 90, 99, 99.9, 100])
    print(f"{name}: blocks {nb} mean {v.mean():.1f} us  p1 {q[0]:.1f} p10 {q[1]:.1f} p50 {q[2]:.1f} p90 {q[3]:.1f} p99 {q[4]:.1f} p99.9 {q[5]:.1f} max {q[6]:.1f}  "
          f"sum/warps(A 4736, B 3552) = {v.sum() / (4736 if name == 'stage A' else 3552) / 1e3:.3f} ms")
