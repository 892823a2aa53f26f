"""Developer check on a B200: parity vs the CPU oracle on small cases + quick timings."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rte_b200  # noqa: E402

pkg = rte_b200.pkg
scene = pkg.scene
import oraclelib  # noqa: E402


def channel_diff(a, b):
    d = np.zeros(a.shape, dtype=np.int32)
    for sh in (0, 8, 16):
        d = np.maximum(d, np.abs(((a >> sh) & 0xFF).astype(np.int32) - ((b >> sh) & 0xFF).astype(np.int32)))
    return d


def compare(r, orc, sc, cam, W, H, flags=0, n_lights=None, **kw):
    r.set_scene(sc, n_lights=n_lights)
    px = r.render(cam, W, H, flags=flags, **kw)
    rows = px.shape[0]
    ids, t = r.hits(rows, W)
    ref = orc.render(sc, cam, W, H, n_lights=n_lights, **kw)
    d = channel_diff(px, ref["pixels"])
    res = {
        "ids_bad": int(np.count_nonzero(ids != ref["ids"])),
        "t_bad": int(np.count_nonzero(t.view(np.uint32) != ref["t"].view(np.uint32))),
        "px_ne": int(np.count_nonzero(d > 0)),
        "px_gt1": int(np.count_nonzero(d > 1)),
        "max": int(d.max()) if d.size else 0,
        "n": int(px.size),
        "hit": int(ref["counts"][3]),
        "cnt": r.counters(),
    }
    return res


def main():
    orc = oraclelib.load("best")
    print("oracle kind:", orc.kind)
    r = pkg.Renderer(0)
    cases = [
        ("R(64,1) 640x480 refcam", scene.reference_scene(64, 1), scene.reference_camera(), 640, 480, {}),
        ("R(64,1) checker 320x240", scene.reference_scene(64, 1, texture=scene.checker_texture(256, 256), sky=scene.checker_texture(512, 256, 32)), scene.reference_camera(), 320, 240, {}),
        ("S(64,2) f0 480x270", None, None, 480, 270, {"scaled": (64, 2, 0)}),
        ("S(64,2) f60 481x271", None, None, 481, 271, {"scaled": (64, 2, 60)}),
        ("R(1024,1) refcam 160x90", scene.reference_scene(1024, 1), scene.reference_camera(), 160, 90, {}),
        ("S(1024,3) f0 384x216", None, None, 384, 216, {"scaled": (1024, 3, 0)}),
        ("S(1024,3) f100 rows", None, None, 3840, 2160, {"scaled": (1024, 3, 100), "kw": dict(y0=5, y1=2160, y_step=97)}),
        ("S(16384,5) f0 96x54", None, None, 96, 54, {"scaled": (16384, 5, 0)}),
        ("S(7000,7) f30 64x36", None, None, 64, 36, {"scaled": (7000, 7, 30)}),
    ]
    for name, sc, cam, W, H, extra in cases:
        if "scaled" in extra:
            n, seed, f = extra["scaled"]
            sc = scene.scaled_scene(n, seed)
            cam = scene.orbit_camera(sc, f)
        kw = extra.get("kw", {})
        for flags in (0, 1):
            t0 = time.time()
            res = compare(r, orc, sc, cam, W, H, flags=flags, **kw)
            c = res.pop("cnt")
            print(f"{name} flags={flags}: {res} exactP={c['exact_primary']} exactS={c['exact_shadow']} ({time.time()-t0:.1f}s)", flush=True)
    # no lights, 1 light, 2 lights
    sc = scene.reference_scene(64, 1)
    for nl in (0, 1, 2):
        res = compare(r, orc, sc, scene.reference_camera(), 200, 150, n_lights=nl)
        res.pop("cnt")
        print(f"R(64,1) n_lights={nl}: {res}")

    # timings
    for (n, seed, W, H) in [(64, 2, 1920, 1080), (1024, 3, 3840, 2160), (16384, 5, 3840, 2160)]:
        sc = scene.scaled_scene(n, seed)
        r.set_scene(sc)
        out = np.empty((H, W), dtype=np.uint32)
        ms_all = []
        for f in range(6):
            cam = scene.orbit_camera(sc, f * 40)
            t0 = time.time()
            r.render(cam, W, H, out=out)
            wall = (time.time() - t0) * 1e3
            ms = r.kernel_ms()
            ms_all.append((wall, ms))
        c = r.counters()
        print(f"S({n},{seed}) {W}x{H}: last wall {ms_all[-1][0]:.2f} ms; kernel ms (prep,primary,shadow) per frame:",
              [[round(v, 3) for v in m[:3]] for _, m in ms_all], "hit", c["hit_pixels"] / c["pixels"],
              "exactP/px", c["exact_primary"] / c["pixels"], "exactS/hit", c["exact_shadow"] / max(1, c["hit_pixels"]), flush=True)


if __name__ == "__main__":
    main()
