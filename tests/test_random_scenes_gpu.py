"""Randomised self-consistency of the CUDA path: for scenes with awkward parameters (zero and large radii, camera or
lights inside spheres, huge light sizes, overlapping spheres, off-axis cameras) the default kernels - per-ray filter,
light cones, warp tiles/beams - must give exactly the frame of the exhaustive mode, which runs the reference's exact
sequence for every ray/sphere pair.  A filter that ever rejected a real hit would show up here."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def random_scene(pkg, seed):
    rng = np.random.default_rng(seed)
    scene = pkg.scene
    n = int(rng.choice([1, 2, 7, 33, 64, 150, 301]))
    extent = float(rng.choice([3.0, 10.0, 40.0]))
    pos = rng.uniform(0, extent, size=(n, 3)).astype(np.float32)
    style = seed % 4
    if style == 0:
        r = rng.uniform(0, 0.99, size=n)            # the reference's radius law
    elif style == 1:
        r = rng.uniform(0, 1.8, size=n)             # effective radius up to 3.2: heavy overlap
    elif style == 2:
        r = np.where(rng.random(n) < 0.3, 0.0, rng.uniform(0.05, 0.7, size=n))   # zero-radius spheres
    else:
        r = rng.uniform(0.3, 0.35, size=n)
    r = r.astype(np.float32)
    spheres = np.concatenate([pos, (r * r)[:, None]], axis=1).astype(np.float32)
    lights = pkg.scene.REFERENCE_LIGHTS.copy()
    k = np.float32(extent / 10.0)
    lights[:, :3] *= k
    lights[:, 3] = np.float32(rng.choice([0.0, 1.0, 20.0, 400.0]))               # light size feeds toEdge
    if seed % 5 == 0:
        lights[0, :3] = pos[0] + np.float32(0.01)                                 # a light inside a sphere
    if seed % 7 == 0:
        lights = lights[:2]
    base = scene.reference_scene(0, 1)
    sc = scene.Scene(spheres=np.ascontiguousarray(spheres), lights=np.ascontiguousarray(lights), texture=base.texture,
                     sky=base.sky, extent=extent, name=f"random{seed}")
    if seed % 3 == 0:                                                            # camera inside the cloud (or a sphere)
        org = tuple(float(v) for v in (pos[n // 2] + np.float32(0.05)))
    else:
        org = tuple(float(v) for v in rng.uniform(-0.5 * extent, 1.5 * extent, size=3))
    cam = scene.Camera(org=org, yaw=float(rng.uniform(0, 360)), pitch=float(rng.uniform(-80, 80)))
    return sc, cam


@pytest.mark.parametrize("seed", list(range(24)))
def test_default_equals_exhaustive_on_random_scenes(seed, renderer, pkg):
    sc, cam = random_scene(pkg, seed)
    renderer.set_scene(sc)
    W, H = 97, 61
    F = pkg.capi
    a = renderer.render(cam, W, H)
    ia, ta = renderer.hits(H, W)
    for flags in (F.ORE_FLAG_EXHAUSTIVE, F.ORE_FLAG_FUSED_SHADOW, F.ORE_FLAG_FUSED_SHADOW | F.ORE_FLAG_EXHAUSTIVE):
        b = renderer.render(cam, W, H, flags=flags)
        ib, tb = renderer.hits(H, W)
        assert np.array_equal(ia, ib), (seed, flags)
        assert np.array_equal(ta.view(np.uint32), tb.view(np.uint32)), (seed, flags)
        assert np.array_equal(a, b), (seed, flags, int(np.count_nonzero(a != b)))


@pytest.mark.parametrize("seed", [1, 4, 10, 15])
def test_random_scenes_match_the_oracle(seed, renderer, oracle_best, pkg):
    sc, cam = random_scene(pkg, seed)
    renderer.set_scene(sc)
    W, H = 64, 40
    px = renderer.render(cam, W, H)
    ids, t = renderer.hits(H, W)
    ref = oracle_best.render(sc, cam, W, H)
    assert np.array_equal(ids, ref["ids"])
    assert np.array_equal(t.view(np.uint32), ref["t"].view(np.uint32))
    d = np.zeros(px.shape, dtype=np.int32)
    for sh in (0, 8, 16):
        d = np.maximum(d, np.abs(((px >> sh) & 255).astype(np.int32) - ((ref["pixels"] >> sh) & 255).astype(np.int32)))
    assert np.count_nonzero(d <= 1) / d.size >= 0.999


@pytest.mark.parametrize("seed", [0, 1, 3, 13, 29])
def test_untame_spheres_do_not_hide_their_cluster_neighbours(seed, renderer, oracle_best, pkg):
    """the shadow sweep walks 32-sphere clusters through one bounding sphere each; a member with huge, infinite or
    NaN coordinates / radius makes its cluster 'always open' instead of poisoning the bound - the frame must still
    equal the exhaustive mode's (and the oracle's)"""
    sc, cam = random_scene(pkg, seed)
    rng = np.random.default_rng(1000 + seed)
    sp = np.repeat(sc.spheres, 2, axis=0)[: max(40, len(sc.spheres))].copy()     # at least two clusters
    sp[:, :3] += rng.uniform(-0.5, 0.5, size=sp[:, :3].shape).astype(np.float32)
    # (none of these is ever hit: the reference's own arithmetic overflows or goes NaN on them)
    weird = [(1e17, 0, 0, 1.0), (0, -3e19, 0, 1e10), (np.inf, 0, 0, 1.0), (-np.inf, np.inf, 0, 2.0), (np.nan, 1, 1, 0.5),
             (1, 1, 1, np.nan), (5e15, 5e15, 5e15, 1e3)]
    for k, w in enumerate(weird):
        sp[(k * 5 + seed) % len(sp)] = np.array(w, dtype=np.float32)
    sc2 = pkg.scene.Scene(spheres=np.ascontiguousarray(sp), lights=sc.lights, texture=sc.texture, sky=sc.sky,
                          extent=sc.extent, name=f"untame{seed}")
    renderer.set_scene(sc2)
    W, H = 72, 44
    a = renderer.render(cam, W, H)
    ia, ta = renderer.hits(H, W)
    b = renderer.render(cam, W, H, flags=pkg.capi.ORE_FLAG_EXHAUSTIVE)
    ib, tb = renderer.hits(H, W)
    assert np.array_equal(ia, ib) and np.array_equal(ta.view(np.uint32), tb.view(np.uint32))
    assert np.array_equal(a, b), int(np.count_nonzero(a != b))
    c = renderer.render(cam, W, H, flags=pkg.capi.ORE_FLAG_FUSED_SHADOW)
    assert np.array_equal(a, c)
    ref = oracle_best.render(sc2, cam, W, H)
    assert np.array_equal(ia, ref["ids"]) and np.array_equal(ta.view(np.uint32), ref["t"].view(np.uint32))
    d = np.zeros(a.shape, dtype=np.int32)
    for sh in (0, 8, 16):
        d = np.maximum(d, np.abs(((a >> sh) & 255).astype(np.int32) - ((ref["pixels"] >> sh) & 255).astype(np.int32)))
    assert np.count_nonzero(d <= 1) >= 0.999 * d.size
