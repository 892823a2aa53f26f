"""Monte-Carlo soundness of the three geometric filters of the shadow sweep, restated in numpy float32 from
csrc/ore_kernels.cuh: per-ray filter, per-light cone, warp beam (also used on cluster bounds).

Truth is float64 geometry: ray i,j (origin S_i, unit direction D_ij) touches the ball (c, R) at some t >= 0.
Whenever that is true, every filter level above the exact test must say "maybe".  Configurations mimic the kernel's:
32 neighbouring origins on a surface patch, 10 sample directions per origin towards a disc-shaped light.
"""
import numpy as np
import pytest

f32 = np.float32
KAPPA = f32(3.814697265625e-06)


def unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def make_group(rng):
    """32 origins on a patch of a sphere surface, 10 directions each towards a light disc (float32 like the kernel)"""
    centre = rng.uniform(-20, 20, 3)
    rad = 10.0 ** rng.uniform(-1, 1)
    n0 = unit(rng.normal(0, 1, 3))
    spread = 10.0 ** rng.uniform(-3, -0.5)
    normals = unit(n0 + spread * rng.normal(0, 1, (32, 3)))
    starts = centre + rad * normals
    light = centre + unit(rng.normal(0, 1, 3) + 1.5 * n0) * 10.0 ** rng.uniform(0.3, 2.2)
    size = 10.0 ** rng.uniform(-2, 1.3)
    jit = rng.uniform(-1, 1, (10, 3)) * size
    dirs = unit((light + jit)[None, :, :] - starts[:, None, :])           # [32,10,3]
    return starts.astype(f32), dirs.astype(f32)


def lane_cones(dirs):
    """cone_of10 + the kernel's margins: axis, ca, sa per lane (float32)"""
    s = dirs.sum(axis=1, dtype=f32)
    inv = (f32(1) / np.sqrt((s * s).sum(axis=1, dtype=f32))).astype(f32)
    ax = (s * inv[:, None]).astype(f32)
    cmin = np.minimum(f32(1), (ax[:, None, :] * dirs).sum(axis=2, dtype=f32).min(axis=1))
    cosa = (cmin - f32(4e-6)).astype(f32)
    sina = (np.sqrt(np.maximum(f32(0), f32(1) - cosa * cosa)) * f32(1.0001) + f32(1e-6)).astype(f32)
    ca = (cosa - f32(0.00196) * sina).astype(f32)
    sa = (f32(1.002) * sina).astype(f32)
    return ax, ca, sa, cmin


def warp_beam(starts, ax, ca, sa):
    b = (starts.sum(axis=0, dtype=f32) / f32(32)).astype(f32)
    e = (starts - b).astype(f32)
    escale = f32(1e-5) * (np.abs(b).sum(dtype=f32) + f32(1))
    s = ax.sum(axis=0, dtype=f32)
    n2 = (s * s).sum(dtype=f32)
    s = (s / np.sqrt(np.maximum(n2, f32(1e-30)))).astype(f32)
    sina = (sa * f32(1 / 1.002)).astype(f32)
    cosa = (ca + f32(0.00196) * sina).astype(f32)
    c1 = np.minimum(f32(1), (ax * s).sum(axis=1, dtype=f32))
    s1 = (np.sqrt(np.maximum(f32(0), f32(1) - c1 * c1)) + f32(1e-6)).astype(f32)
    cw = (c1 * cosa - s1 * sina - f32(2e-6)).astype(f32).min()
    ai = (e * s).sum(axis=1, dtype=f32)
    p = (e - ai[:, None] * s).astype(f32)
    rp = np.sqrt((p * p).sum(axis=1, dtype=f32)).max()
    amin = ai.min()
    if not (n2 > 1e-12 and cw > 0.3):
        return None                                           # wforce: no warp-level culling
    sinw = f32(np.sqrt(max(f32(0), f32(1) - cw * cw)) * f32(1.0001) + f32(1e-6))
    return dict(b=b, s=s, tan=f32(sinw / cw * f32(1.0001)), k1=f32(-(amin - escale)), k2=f32(rp * f32(1.0001) + escale))


def beam_may_touch(beam, c, Rp):
    if beam is None:
        return np.ones(len(c), dtype=bool)
    L = (beam["b"] - c).astype(f32)
    LL = (L * L).sum(axis=1, dtype=f32)
    Rq = (Rp * f32(1.0001) + LL * f32(1e-12) + f32(1e-6)).astype(f32)
    slack = (LL * f32(2e-6)).astype(f32)
    sc = (-(L * beam["s"]).sum(axis=1, dtype=f32)).astype(f32)
    u = (sc + Rq + beam["k1"]).astype(f32)
    thr = (u * beam["tan"] + Rq + beam["k2"]).astype(f32)
    d2 = (LL - sc * sc).astype(f32)
    return (u >= 0) & (d2 <= thr * thr + slack)


def lane_filters(starts, dirs, ax, ca, sa, c, Rp):
    """per lane and sphere: cone test (2.3) and per-ray filter (2.1); returns [32,n] cone pass, [32,10,n] ray pass"""
    l = (starts[:, None, :] - c[None, :, :]).astype(f32)                       # [32,n,3]
    LL = (l * l).sum(axis=2, dtype=f32)
    Cm = (LL * (f32(1) - KAPPA) - (Rp * Rp)[None, :]).astype(f32)
    sv = np.where(Cm > 1e-20, np.sqrt(np.maximum(Cm, f32(0))), f32(-3e38)).astype(f32)
    T = (ca[:, None] * sv - sa[:, None] * Rp[None, :]).astype(f32)
    cone = ((ax[:, None, :] * l).sum(axis=2, dtype=f32) + T) < 0
    ray = ((dirs[:, :, None, :] * l[:, None, :, :]).sum(axis=3, dtype=f32) + sv[:, None, :]) < 0
    return cone, ray


def truth(starts, dirs, c, R):
    S = starts.astype(np.float64)[:, None, None, :]
    D = dirs.astype(np.float64)[:, :, None, :]
    l = S - c.astype(np.float64)[None, None, :, :]                              # [32,1,n,3]
    b = (D * l).sum(axis=3)
    cc = (l * l).sum(axis=3) - (R.astype(np.float64) ** 2)[None, None, :]
    disc = b * b - cc
    return (disc >= 0) & ((-b + np.sqrt(np.maximum(disc, 0))) >= 0)             # far root >= 0  [32,10,n]


@pytest.mark.parametrize("seed", range(6))
def test_no_filter_level_rejects_a_geometric_hit(seed):
    rng = np.random.default_rng(500 + seed)
    hits = misses = culled = 0
    for _ in range(60):
        starts, dirs = make_group(rng)
        ax, ca, sa, cmin = lane_cones(dirs)
        if not np.all(cmin > 0.5):
            continue                                                             # the kernel forces "maybe" there
        beam = warp_beam(starts, ax, ca, sa)
        n = 400
        # spheres: half placed along the rays (so that many are really hit), half anywhere
        i, j = rng.integers(0, 32, n), rng.integers(0, 10, n)
        t = 10.0 ** rng.uniform(-2, 2.5, n)
        R = (10.0 ** rng.uniform(-2, 1, n)).astype(f32)
        c = starts[i].astype(np.float64) + dirs[i, j].astype(np.float64) * t[:, None]
        c += unit(rng.normal(0, 1, (n, 3))) * (R * rng.uniform(0, 2.0, n))[:, None]
        c[n // 2:] = starts.mean(axis=0) + rng.normal(0, 1, (n - n // 2, 3)) * 10.0 ** rng.uniform(0, 2)
        c = c.astype(f32)
        Rp = np.nextafter((R * f32(1.000004)).astype(f32), f32(np.inf))          # R'^2 >= (1+kappa) R^2
        hit = truth(starts, dirs, c, R)                                          # [32,10,n]
        cone, ray = lane_filters(starts, dirs, ax, ca, sa, c, Rp)
        wb = beam_may_touch(beam, c, Rp)
        assert not np.any(hit & ~ray), "per-ray filter rejected a hit"
        assert not np.any(hit.any(axis=1) & ~cone), "light cone rejected a hit"
        assert not np.any(hit.any(axis=(0, 1)) & ~wb), "warp beam rejected a hit"
        # a cluster bound = any ball containing the member: the beam test must be monotone under containment
        grow = (Rp * f32(rng.uniform(1.0, 3.0))).astype(f32)
        assert not np.any(wb & ~beam_may_touch(beam, c, grow)), "beam test is not monotone in the radius"
        any_hit = hit.any(axis=(0, 1))
        hits += int(any_hit.sum())
        misses += int((~any_hit).sum())
        culled += int((~any_hit & ~wb).sum())
    assert hits > 500, "the sample must contain real hits"
    assert culled > 0.6 * misses, "and the beam test must actually cull (it removes ~90 % of the misses here)"
