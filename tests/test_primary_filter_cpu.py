"""Monte-Carlo soundness of the primary-ray filters restated in numpy from csrc: the per-pixel
one-FFMA filter on the per-frame sphere records of prep_frame_kernel, and the 32x8 tile cone.

Truth is float64 geometry on the reference's own ray directions (kernel.cu:1624-1631, camera::rotateDir :252-255):
positive discriminant and far root >= 0.  A real hit must pass the pixel filter, and a tile containing one must pass
the tile cone.
"""
import math

import numpy as np
import pytest

f32 = np.float32
KP = 7.62939453125e-06          # ORE_KAPPA_PRIMARY = 2^-17
P = 8                           # tile rows


def frame_tables(W, H, aspect):
    x = np.arange(W)
    y = np.arange(H)
    dx = (np.float64(aspect) * (2 * (x + 0.5) / np.float64(f32(W))) - 1).astype(f32)
    hw = f32(f32(H) / f32(W))
    dy = (np.float64(aspect) * (2 * (y + 0.5) / np.float64(f32(H))) * np.float64(hw) - 1).astype(f32)
    return dx, dy


def directions(dx, dy, fz, cp, sp, cy, sy):
    vx, vy = np.meshgrid(dx, dy)                                  # [H,W]
    vz = np.full_like(vx, fz)
    ln = np.sqrt((vx * vx + vy * vy + vz * vz).astype(f32)).astype(f32)
    nx, ny, nz = (vx / ln).astype(f32), (vy / ln).astype(f32), (vz / ln).astype(f32)
    y = (ny * cp - nz * sp).astype(f32)
    z = (ny * sp + nz * cp).astype(f32)
    x = (nx * cy + z * sy).astype(f32)
    z2 = (-nx * sy + z * cy).astype(f32)
    return np.stack([x, y, z2], axis=-1)


@pytest.mark.parametrize("seed", range(5))
def test_primary_filters_never_reject_a_hit(seed):
    rng = np.random.default_rng(900 + seed)
    W, H = 256, 128
    aspect = f32(math.tan(90 * 0.5 * 3.1415 / 180))
    ez = f32(-1.0) / aspect
    fz = f32(0.0) - ez
    dx, dy = frame_tables(W, H, aspect)
    delta = 2.0 * float(aspect) / W
    hx, hy = 16.0 * delta, (0.5 * (P - 1) + 0.5) * delta
    xr = math.sqrt(hx * hx + hy * hy) / abs(float(fz))
    a = math.asin(xr) * 1.001 + 1e-6
    tile_ca, tile_sa = f32(math.cos(a) * (1 - 1e-6) - 1e-6), f32(math.sin(a) * (1 + 1e-6) + 1e-6)
    n_hit = n_tiles_culled = n_tiles = 0
    for _ in range(6):
        yaw, pitch = rng.uniform(0, 360), rng.uniform(-80, 80)
        cp, sp = f32(math.cos(math.radians(pitch))), f32(math.sin(math.radians(pitch)))
        cy, sy = f32(math.cos(math.radians(yaw))), f32(math.sin(math.radians(yaw)))
        O = rng.uniform(-10, 10, 3).astype(f32)
        D = directions(dx, dy, fz, cp, sp, cy, sy).astype(np.float64)            # [H,W,3]
        n = 60
        # spheres: most of them in front of the camera along some pixel's ray
        py, px = rng.integers(0, H, n), rng.integers(0, W, n)
        t = 10.0 ** rng.uniform(-0.5, 2.5, n)
        member = (10.0 ** rng.uniform(-1.5, 0.7, n)).astype(f32)                 # effective radius
        c = O.astype(np.float64) + D[py, px] * t[:, None] + rng.normal(0, 1, (n, 3)) * member[:, None] * 1.5
        c[n - 10:] = O + rng.normal(0, 1, (10, 3)) * 30                           # anywhere, incl. behind
        c = c.astype(f32)
        # per-frame records (prep_frame_kernel, double precision)
        L = (O[None, :] - c).astype(f32).astype(np.float64)                       # reference forms L in float
        LL = (L * L).sum(axis=1)
        r4 = (member * member).astype(f32).astype(np.float64)
        Cm = LL * (1 - KP) - r4 * (1 + KP)
        always = ~((Cm > 1e-9 * LL) & (Cm > 1e-30))
        sv = np.sqrt(np.where(always, 1.0, Cm))
        cpd, spd, cyd, syd = float(cp), float(sp), float(cy), float(sy)
        Mx = cyd * L[:, 0] - syd * L[:, 2]
        My = spd * syd * L[:, 0] + cpd * L[:, 1] + spd * cyd * L[:, 2]
        Mz = cpd * syd * L[:, 0] - spd * L[:, 1] + cpd * cyd * L[:, 2]
        ra, rb, rc = (Mx / sv).astype(f32), (My / sv).astype(f32), (float(fz) * Mz / sv).astype(f32)
        Rpp = np.sqrt(np.maximum(LL - Cm, 0))
        Wd = (float(tile_ca) * sv - float(tile_sa) * Rpp - 4e-6 * np.sqrt(LL) - 1e-30).astype(f32)
        # truth per pixel and sphere
        b = np.einsum("hwk,nk->hwn", D, L)
        disc = b * b - (LL - r4)[None, None, :]
        hit = (disc >= 0) & ((-b + np.sqrt(np.maximum(disc, 0))) >= 0)            # [H,W,n]
        # per-pixel filter: fma(dy, b', fma(dx, a', c')) <= -|v| (1 - 2^-20)
        nv = np.sqrt((dx[None, :] ** 2 + dy[:, None] ** 2 + fz * fz).astype(f32)).astype(f32)
        lhs = (dy[:, None, None] * rb[None, None, :] + (dx[None, :, None] * ra[None, None, :] + rc[None, None, :])).astype(f32)
        passed = (lhs <= (-nv * f32(0.99999905))[:, :, None]) | always[None, None, :]
        assert not np.any(hit & ~passed), "pixel filter rejected a hit"
        # tile cone
        for ty in range(H // P):
            cyt = f32(0.5) * (dy[ty * P] + dy[ty * P + P - 1])
            for tx in range(W // 32):
                cxt = dx[tx * 32] + f32(15.5) * f32(delta)
                inv = f32(1) / np.sqrt(f32(cxt * cxt + cyt * cyt + fz * fz))
                axv = np.array([cxt * inv, cyt * inv, fz * inv], dtype=f32)
                hA = (axv[0] * Mx.astype(f32) + (axv[1] * My.astype(f32) + (axv[2] * Mz.astype(f32) + Wd))).astype(f32)
                cone_ok = (hA <= 0) | always
                th = hit[ty * P:(ty + 1) * P, tx * 32:(tx + 1) * 32].any(axis=(0, 1))
                assert not np.any(th & ~cone_ok), "tile cone rejected a tile with a hit"
                n_tiles += n
                n_tiles_culled += int((~cone_ok).sum())
        n_hit += int(hit.any(axis=(0, 1)).sum())
    assert n_hit > 100
    assert n_tiles_culled > 0.5 * n_tiles, "the tile cone must cull most (tile, sphere) pairs"
