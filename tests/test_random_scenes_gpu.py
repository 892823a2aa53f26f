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
    for flags in (F.ORE_FLAG_EXHAUSTIVE, F.ORE_FLAG_NO_WARP_CULL, F.ORE_FLAG_PER_RAY_SHADOW,
                  F.ORE_FLAG_PER_RAY_SHADOW | F.ORE_FLAG_EXHAUSTIVE, F.ORE_FLAG_FUSED_SHADOW):
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
