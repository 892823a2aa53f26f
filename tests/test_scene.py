import numpy as np


def test_lcg_is_msvc_rand(pkg):
    rng = pkg.scene.Lcg(1)
    assert [rng.next() for _ in range(5)] == [41, 18467, 6334, 26500, 19169]   # MSVC rand() after srand(1)


def test_reference_scene_formula(pkg):
    sc = pkg.scene.reference_scene(64, 1)
    assert sc.spheres.shape == (64, 4) and sc.spheres.dtype == np.float32
    # kernel.cu:1190: centre = (rand()%100)/10, r = (rand()%100)/100, member = r*r
    assert sc.spheres[0].tolist() == [np.float32(41 % 100) / np.float32(10), np.float32(18467 % 100) / np.float32(10),
                                      np.float32(6334 % 100) / np.float32(10), (np.float32(0.0)) ** 2]
    assert float(sc.spheres[:, :3].max()) <= 9.9 and float(sc.spheres[:, 3].max()) <= 0.99 ** 2 + 1e-6
    assert np.array_equal(sc.lights, pkg.scene.REFERENCE_LIGHTS)
    assert abs(float(sc.aspect) - 0.999953687) < 1e-6


def test_scaled_scene_density(pkg):
    sc = pkg.scene.scaled_scene(1024, 3)
    assert abs(sc.extent - 10 * 16 ** (1 / 3)) < 1e-9
    assert float(sc.spheres[:, :3].max()) <= sc.extent
    cam = pkg.scene.orbit_camera(sc, 0)
    assert cam.yaw == 180.0


def test_sprite_format(pkg):
    t = pkg.scene.smooth_texture(64, 32, 5)
    assert t.r.shape == (64 * 32,) and t.r.dtype == np.float32
    # plane = byte/255 (Sprite.cpp:44-46)
    b = np.rint(t.r * 255).astype(np.int32)
    assert np.array_equal((b.astype(np.float32) / np.float32(255)), t.r)
