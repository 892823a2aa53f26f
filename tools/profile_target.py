"""Short program for ncu: a few frames of one workload through the C ABI (no oracle, no CPU work)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import rte_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="8k1024")
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--json", default=None, help="write the per-frame device counters here")
    a = ap.parse_args()
    pkg = rte_b200.pkg
    W, H, sc, camera, desc = bench.make_workload(pkg, a.workload)
    r = pkg.Renderer(0)
    r.set_scene(sc)
    out = np.empty((H, W), dtype=np.uint32)
    frames = []
    for f in range(a.frames):
        r.render(camera(f), W, H, out=out, flags=a.flags)
        c = r.counters()
        print(f, [round(v, 3) for v in r.kernel_ms()[:3]], c["hit_pixels"], "L1/warp", round(c["beam_l1"] / max(1, c["hit_pixels"] / 32), 1), "L2/px", round(c["beam_l2"] / max(1, c["hit_pixels"]), 2), "exactS/px", round(c["exact_shadow"] / max(1, c["hit_pixels"]), 1), "exactP/px", round(c["exact_primary"] / c["pixels"], 2),
              "sweep steps/blk", round(c["sweep_steps"] / max(1, c["hit_pixels"] / 32), 1), "primary steps/tile", round(c["primary_steps"] / max(1, c["pixels"] / 256), 1), flush=True)
        frames.append(dict(c, kernel_ms=r.kernel_ms()))
    if a.json:
        import json
        json.dump({"workload": a.workload, "frames": frames}, open(a.json, "w"), indent=1)
    r.close()


if __name__ == "__main__":
    main()
