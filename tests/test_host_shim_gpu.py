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


def test_headless_harness_with_an_obj_mesh(renderer, pkg, tmp_path):
    """onStart() loading an OBJ through the shim's own loader / BVH builder == the Python binding fed with the
    fixture the reference's loader produced"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases

    subprocess.check_call(["make", "-s", "-C", HOST])
    obj = tmp_path / "torus.obj"
    pkg.scene.write_torus_obj(str(obj))
    out = tmp_path / "frame.ppm"
    W, H, frames = 256, 192, 1
    subprocess.run([os.path.join(HOST, "ore_headless"), str(W), str(H), str(frames), "64", str(out), str(obj)],
                   capture_output=True, text=True, check=True)
    data = out.read_bytes()
    header = b"P6\n%d %d\n255\n" % (W, H)
    rgb = np.frombuffer(data[len(header):], dtype=np.uint8).reshape(H, W, 3)[::-1]
    got = (rgb[..., 0].astype(np.uint32) << 16) | (rgb[..., 1].astype(np.uint32) << 8) | rgb[..., 2]
    sc = pkg.scene.reference_scene(64, 1)
    sc.mesh = cases.torus_mesh()
    yaw, pitch = 180.0, 15.0
    yr, pr = math.radians(yaw), math.radians(pitch)
    org = tuple(float(np.float32(v)) for v in (5 - 12 * math.cos(pr) * math.sin(yr), 5 + 12 * math.sin(pr),
                                               5 - 12 * math.cos(pr) * math.cos(yr)))
    renderer.set_scene(sc)
    want = renderer.render(pkg.scene.Camera(org=org, yaw=yaw, pitch=pitch), W, H)
    renderer.set_mesh(None)
    assert np.count_nonzero(got == want) / want.size >= 0.999
    ids, _ = renderer.hits(H, W)


def test_sprite_loads_ppm_and_bmp_into_the_reference_plane_format(tmp_path):
    """sprite(file): planar float r,g,b = byte/255, row-major from the top row (Sprite.cpp:28-52), no OpenCV"""
    import ctypes as C
    import struct

    subprocess.check_call(["make", "-s", "-C", HOST])
    # the shim leaves the window callbacks (getScreenWidth, ...) to the window layer: bind lazily
    lib = C.CDLL(os.path.join(HOST, "libore_host.so"), mode=os.RTLD_LAZY)
    fp = C.POINTER(C.c_float)
    lib.ore_host_sprite_load.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), fp, fp, fp, C.c_int]
    rng = np.random.default_rng(3)
    w, h = 37, 21                                    # odd width: BMP rows are padded to 4 bytes
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)      # R,G,B, row 0 = top
    ppm = tmp_path / "t.ppm"
    ppm.write_bytes(b"P6\n# comment\n%d %d\n255\n" % (w, h) + img.tobytes())
    stride = (w * 3 + 3) & ~3
    rows = b"".join(img[y, :, ::-1].tobytes() + b"\0" * (stride - w * 3) for y in range(h - 1, -1, -1))  # bottom-up BGR
    bmp = tmp_path / "t.bmp"
    bmp.write_bytes(b"BM" + struct.pack("<IHHI", 54 + len(rows), 0, 0, 54) +
                    struct.pack("<IiiHHIIiiII", 40, w, h, 1, 24, 0, len(rows), 2835, 2835, 0, 0) + rows)
    want = [np.ascontiguousarray((img[:, :, c].astype(np.float32) / np.float32(255)).reshape(-1)) for c in range(3)]
    for path in (ppm, bmp):
        ww, hh = C.c_int(0), C.c_int(0)
        planes = [np.zeros(w * h, dtype=np.float32) for _ in range(3)]
        rc = lib.ore_host_sprite_load(str(path).encode(), C.byref(ww), C.byref(hh),
                                      *[p.ctypes.data_as(fp) for p in planes], w * h)
        assert rc == 0 and (ww.value, hh.value) == (w, h), path
        for got, exp in zip(planes, want):
            assert np.array_equal(got, exp), path


@pytest.mark.gpu
def test_headless_present_by_pointer_shows_the_same_frames(tmp_path):
    """present = pointer (the window keeps the pointer update() hands it instead of memcpy'ing the frame) ends on the
    same last frame as the reference-style copy, synchronous and pipelined"""
    outs = []
    for pipelined, present in ((0, "copy"), (1, "copy"), (1, "pointer"), (0, "pointer")):
        out = tmp_path / f"p{pipelined}_{present}.ppm"
        res = subprocess.run([os.path.join(HOST, "ore_headless"), "320", "200", "5", "64", str(out), "-", str(pipelined), present],
                             capture_output=True, text=True, timeout=120)
        assert res.returncode == 0, res.stderr
        outs.append(out.read_bytes())
    assert outs[0] == outs[1] == outs[2] == outs[3]
