"""The C++ drop-in shims (onStart/update/memManager/sprite over the C ABI) driven by the headless window."""
import json
import math
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "ray-tracer-engine_b200", "host")


def test_headless_harness_matches_the_python_binding(renderer, pkg, tmp_path):
    subprocess.check_call(["make", "-s", "-C", HOST])
    out = tmp_path / "frame.ppm"
    W, H, frames = 320, 240, 2
    res = subprocess.run([os.path.join(HOST, "ore_headless"), str(W), str(H), str(frames), "64", str(out)],
                         capture_output=True, text=True, check=True)
    info = json.loads(res.stdout.strip().splitlines()[-1])
    assert info["width"] == W and info["frames"] == frames
    data = out.read_bytes()
    header = b"P6\n%d %d\n255\n" % (W, H)
    assert data.startswith(header)
    rgb = np.frombuffer(data[len(header):], dtype=np.uint8).reshape(H, W, 3)[::-1]   # PPM is top-down
    got = (rgb[..., 0].astype(np.uint32) << 16) | (rgb[..., 1].astype(np.uint32) << 8) | rgb[..., 2]

    # same scene through the Python binding: reference generator R(64,1), same procedural textures, last orbit frame
    sc = pkg.scene.reference_scene(64, 1)
    f = frames - 1
    yaw, pitch = 180.0 + 360.0 * f / frames, 15.0
    yr, pr = math.radians(yaw), math.radians(pitch)
    org = tuple(float(np.float32(v)) for v in (5 - 12 * math.cos(pr) * math.sin(yr), 5 + 12 * math.sin(pr),
                                               5 - 12 * math.cos(pr) * math.cos(yr)))
    cam = pkg.scene.Camera(org=org, yaw=float(np.float32(yaw)), pitch=float(np.float32(pitch)))
    renderer.set_scene(sc)
    want = renderer.render(cam, W, H)
    same = np.count_nonzero(got == want) / want.size
    assert same >= 0.999, f"only {same:.5f} of the harness frame equals the binding's frame"
    assert info["checksum"] == int(got.astype(np.uint64).sum())
