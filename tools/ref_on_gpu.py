"""Run the reference's own CUDA kernel (built for sm_100) next to ours on the same scene: parity + timing."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import oraclelib  # noqa: E402
import rte_b200  # noqa: E402


def chdiff(a, b):
    d = np.zeros(a.shape, dtype=np.int32)
    for sh in (0, 8, 16):
        d = np.maximum(d, np.abs(((a >> sh) & 255).astype(np.int32) - ((b >> sh) & 255).astype(np.int32)))
    return d


def main():
    pkg = rte_b200.pkg
    out = {}
    for wl, frames in (("1080p64", 6), ("4k1024", 3)):
        W, H, sc, camera, desc = bench.make_workload(pkg, wl)
        cams = [camera(i * 7) for i in range(frames)]
        r = pkg.Renderer(0)
        r.set_scene(sc)
        mine = r.render(cams[-1], W, H)
        res = {"workload": desc}
        for fast in (False, True):
            ref = oraclelib.RefGpu(fast=fast)
            px, ms_update, ms_kernel = ref.render(sc, cams, W, H)
            d = chdiff(mine, px)
            res["fast_math" if fast else "default_flags"] = {
                "ms_per_update": ms_update, "ms_per_kernel": ms_kernel,
                "mrays_update": W * H / ms_update / 1e3, "mrays_kernel": W * H / ms_kernel / 1e3,
                "pixels_within_1lsb_of_ours": float(np.count_nonzero(d <= 1)) / d.size,
                "pixels_identical_to_ours": float(np.count_nonzero(d == 0)) / d.size, "max_diff": int(d.max())}
        r.close()
        out[wl] = res
        print(json.dumps({wl: res}), flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ref_on_gpu.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
