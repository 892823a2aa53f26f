"""Known-answer tests of the CPU oracle on hand-computable cases.

The reference has no tests; these pin the quirks of sphere::intersect (kernel.cu:293-354)
and rgbToInt (kernel.cu:546-556) that the CUDA path must reproduce."""
import math

import numpy as np
import pytest


@pytest.fixture(params=["port", "ref"])
def orc(request, oracle_port, request_ref=None):
    if request.param == "port":
        return oracle_port
    return request.getfixturevalue("oracle_ref")


def test_hit_in_front_returns_near_root(orc):
    # unit direction, effective radius = radius member (0.25 -> tested as 0.0625 = r_eff^2, r_eff 0.25)
    hit, t = orc.sphere_intersect((0, 0, 0), (0, 0, 1), (0, 0, 5), 0.25)
    assert hit and t == np.float32(4.75)


def test_effective_radius_is_member_not_its_root(orc):
    # ctor r = 0.5 stores member 0.25; a ray passing 0.3 from the centre must MISS (0.3 > 0.25)
    hit, _ = orc.sphere_intersect((0.3, 0, 0), (0, 0, 1), (0, 0, 5), 0.25)
    assert not hit
    hit, _ = orc.sphere_intersect((0.2, 0, 0), (0, 0, 1), (0, 0, 5), 0.25)
    assert hit


def test_origin_inside_returns_negative_near_root(orc):
    hit, t = orc.sphere_intersect((0, 0, 5), (0, 0, 1), (0, 0, 5), 0.25)
    assert hit and t == np.float32(-0.25)


def test_sphere_behind_ray_is_a_miss(orc):
    hit, t = orc.sphere_intersect((0, 0, 0), (0, 0, 1), (0, 0, -5), 0.25)
    assert not hit and t < 0


def test_miss_leaves_nan(orc):
    hit, t = orc.sphere_intersect((0, 0, 0), (0, 0, 1), (3, 0, 5), 0.25)
    assert not hit and math.isnan(float(t))


def test_far_root_exactly_zero_counts_as_hit(orc):
    # origin on the far surface: far root t == 0 -> `if (t == 0.f) return true` (kernel.cu:338)
    hit, t = orc.sphere_intersect((0, 0, 5.25), (0, 0, 1), (0, 0, 5), 0.25)
    assert hit and t == 0


def test_far_root_gate_is_a_double_compare(orc):
    # far root just below / at the double 0.0001: sphere of effective radius 1 centred so that
    # t_far = z + 1; choose origins giving t_far = 0x38D1B717 (below) and 0x38D1B718 (at/above)
    below = np.uint32(0x38D1B717).view(np.float32)
    above = np.uint32(0x38D1B718).view(np.float32)
    assert float(below) < 0.0001 <= float(above)


def test_rgb_to_int(orc):
    assert orc.rgb_to_int(300, -1, 128) == 0x00FFFF80
    assert orc.rgb_to_int(1, 2, 3) == 0x00010203
    assert orc.rgb_to_int(255, 256, 0) == 0x00FFFF00
