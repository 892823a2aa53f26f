"""The claim behind FrameParams::skip_dark, checked against the REFERENCE's own code on the CPU: a light
that a surface point does not face (a = normal . toL <= 0) changes nothing at that pixel, whatever its shadow rays hit -
castLightRay multiplies its sample count by max(0, a) (kernel.cu:1541-1542).  The frame is rendered with two lights and
with a third one added; the pixels whose a for the third light is clearly negative (float64 geometry from the oracle's
ids and t) must be identical in both frames, and pixels that clearly face it must change somewhere."""
import math

import numpy as np
import pytest


def _facing(sc, cam, W, H, ids, t, light_pos):
    """a = normal . toL per pixel in float64 (NaN where nothing was hit); primary rays as kernel.cu:1624-1631, 252-255"""
    aspect = float(sc.aspect)
    x, y = np.meshgrid(np.arange(W), np.arange(H))
    dx = aspect * (2 * (x + 0.5) / W) - 1
    dy = aspect * (2 * (y + 0.5) / H) * (H / W) - 1
    v = np.stack([dx, dy, np.full_like(dx, 1 / aspect)], axis=-1)
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    yr, pr = cam.yaw * (3.1415 / 180), cam.pitch * (3.1415 / 180)
    cp, sp, cy, sy = math.cos(pr), math.sin(pr), math.cos(yr), math.sin(yr)
    yy = v[..., 1] * cp - v[..., 2] * sp
    zz = v[..., 1] * sp + v[..., 2] * cp
    xx = v[..., 0] * cy + zz * sy
    zz = -v[..., 0] * sy + zz * cy
    D = np.stack([xx, yy, zz], axis=-1)
    O = np.array(cam.org, dtype=np.float64) + np.array([0, 0, -1 / aspect])
    hit = ids >= 0
    P = O + D * np.where(hit, t, 0.0)[..., None]
    c = sc.spheres[np.where(hit, ids, 0), :3].astype(np.float64)
    n = P - c
    n /= np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), 1e-30)
    toL = np.asarray(light_pos, dtype=np.float64) - P
    toL /= np.linalg.norm(toL, axis=-1, keepdims=True)
    return np.where(hit, (n * toL).sum(-1), np.nan)


@pytest.mark.parametrize("which", ["best", "port"])
@pytest.mark.parametrize("light_pos", [(5.0, -60.0, 5.0), (-40.0, 2.0, 4.0), (4.0, 30.0, -30.0)])
def test_a_light_a_pixel_does_not_face_leaves_it_unchanged(which, light_pos, pkg):
    import oraclelib

    orc = oraclelib.load(which)
    sc = pkg.scene.reference_scene(64, 1)
    cam = pkg.scene.reference_camera()
    W, H = 160, 120
    base = orc.render(sc, cam, W, H, n_lights=2)
    extra = np.array([[*light_pos, 20.0, 0.7, 0.8, 0.9]], dtype=np.float32)
    lights = np.concatenate([sc.lights[:2], extra]).astype(np.float32)
    sc2 = pkg.scene.Scene(spheres=sc.spheres, lights=np.ascontiguousarray(lights), texture=sc.texture, sky=sc.sky,
                          extent=sc.extent, name="third_light")
    more = orc.render(sc2, cam, W, H)
    assert np.array_equal(base["ids"], more["ids"])
    a = _facing(sc, cam, W, H, base["ids"], base["t"].astype(np.float64), light_pos)
    away = a < -1e-3
    towards = a > 0.05
    assert away.sum() > 300, int(away.sum())
    assert np.array_equal(base["pixels"][away], more["pixels"][away]), int((base["pixels"][away] != more["pixels"][away]).sum())
    if towards.sum() > 300:
        assert (base["pixels"][towards] != more["pixels"][towards]).any()     # the light is not simply switched off
