"""Kernel times for a full frame vs 1/P of its rows on ONE GPU (isolates partial-frame effects from NVLink effects)."""
import os
import sys
import statistics

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import rte_b200  # noqa: E402

pkg = rte_b200.pkg
W, H, sc, camera, desc = bench.make_workload(pkg, "8k1024")
r = pkg.Renderer(0)
r.set_scene(sc)
P = 8
for name, kw in (("full", {}), ("block-interleaved 1/8", dict(y0=0, y1=H, y_step=8 * P, y_block=8)),
                 ("row-interleaved 1/8", dict(y0=0, y1=H, y_step=P)), ("contiguous 1/8 (middle)", dict(y0=H // 2, y1=H // 2 + H // P))):
    ms = []
    for f in range(6):
        out = r.render(camera(f), W, H, **kw)
        if f >= 2:
            ms.append(r.kernel_ms()[:3])
    c = r.counters()
    print(f"{name:28s} rows {out.shape[0]:5d} hit {c['hit_pixels']:8d} prep/primary/shadow ms:",
          [round(statistics.mean(m[i] for m in ms), 3) for i in range(3)], flush=True)
r.close()
