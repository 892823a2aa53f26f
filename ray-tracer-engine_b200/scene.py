"""Synthetic sphere scenes, camera orbit and sprite-format textures for the render path.

The reference ships no assets and no scene files: its spheres come from unseeded
`rand()` (`/root/reference/kernel.cu:1189-1191`), its lights and camera are literals
(`kernel.cu:1695,1708-1712`, `:261`) and its textures are image files on the author's
disk (`:1700,1706`).  SURVEY.md section 8(d) therefore fixes harness-defined inputs; a
scene is an INPUT to both the CUDA path and the CPU oracle, so parity never depends on
these choices.

Everything here is plain numpy on the host (float32 where the reference stores float).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# kernel.cu:1701  float aspect = tan((90 * 0.5 * 3.1415) / 180);   (double tan, stored as float)
REFERENCE_ASPECT = np.float32(math.tan((90 * 0.5 * 3.1415) / 180))
# kernel.cu:1708-1710  pos.xyz, size, r, g, b
REFERENCE_LIGHTS = np.array(
    [[20, 20, 20, 20, 1, 0, 0], [0, 20, -20, 20, 0, 0, 1], [0, 20, 0, 20, 0, 1, 0]], dtype=np.float32
)
SKY_SIZE = 10000.0  # kernel.cu:1700


class Lcg:
    """MSVC-compatible `rand()`: s = s*214013 + 2531011; out = (s >> 16) & 0x7fff.

    The reference calls unseeded `rand()` (seed 1) on Windows (`kernel.cu:1190`).
    """

    def __init__(self, seed: int = 1):
        self.s = seed & 0xFFFFFFFF

    def next(self) -> int:
        self.s = (self.s * 214013 + 2531011) & 0xFFFFFFFF
        return (self.s >> 16) & 0x7FFF


@dataclass
class Camera:
    """camera::Org, Camyaw, Campitch (degrees) - kernel.cu:237-262,1695."""

    org: tuple = (4.0, 3.0, 10.0)
    yaw: float = 180.0
    pitch: float = -20.0


@dataclass
class Sprite:
    """sprite format (sprite.h:11-47, Sprite.cpp:28-52): planar float r,g,b = byte/255, row-major."""

    width: int
    height: int
    r: np.ndarray
    g: np.ndarray
    b: np.ndarray

    @staticmethod
    def from_bytes(rgb: np.ndarray) -> "Sprite":
        """rgb: uint8 [h, w, 3] in R,G,B order; planes are (float)byte / 255 (Sprite.cpp:44-46)."""
        h, w, _ = rgb.shape
        planes = [np.ascontiguousarray((rgb[:, :, c].astype(np.float32) / np.float32(255)).reshape(-1)) for c in range(3)]
        return Sprite(w, h, planes[0], planes[1], planes[2])


@dataclass
class Mesh:
    """A triangle mesh WITH its flat BVH, as the reference's `mesh` holds it (kernel.cu:559-1017): the arrays the
    reference's own OBJ loader and createBvhMesh() produce; this path consumes them, it does not rebuild them."""

    tris: np.ndarray          # [n,27] float32: points[3], normal, vecNormal[3], vt[3] (kernel.cu:206-212)
    has_normals: bool
    box_bounds: np.ndarray    # [b,6] float32: leaf cube bounds[0], bounds[1]
    box_offsets: np.ndarray   # [b+1] int32
    box_indices: np.ndarray   # [sum] int32 triangle indices per leaf, in leaf order

    @property
    def n_tris(self) -> int:
        return int(self.tris.shape[0])

    @property
    def n_boxes(self) -> int:
        return int(self.box_bounds.shape[0])

    @staticmethod
    def from_arrays(d) -> "Mesh":
        return Mesh(np.ascontiguousarray(d["tris"], dtype=np.float32), bool(d["has_normals"]),
                    np.ascontiguousarray(d["box_bounds"], dtype=np.float32),
                    np.ascontiguousarray(d["box_offsets"], dtype=np.int32),
                    np.ascontiguousarray(d["box_indices"], dtype=np.int32))


def write_torus_obj(path: str, centre=(5.0, 5.0, 5.0), major=2.5, minor=0.9, nu=24, nv=12) -> int:
    """A torus as `v / vt / vn / f a/b/c` triangles - the OBJ dialect the reference's loader parses
    (kernel.cu:594-700).  Returns the triangle count."""
    verts, uvs, norms, faces = [], [], [], []
    for i in range(nu):
        a = 2 * math.pi * i / nu
        for j in range(nv):
            b = 2 * math.pi * j / nv
            nx, ny, nz = math.cos(a) * math.cos(b), math.sin(b), math.sin(a) * math.cos(b)
            verts.append((centre[0] + (major + minor * math.cos(b)) * math.cos(a), centre[1] + minor * math.sin(b),
                          centre[2] + (major + minor * math.cos(b)) * math.sin(a)))
            norms.append((nx, ny, nz))
            uvs.append((i / nu, j / nv))
    def vid(i, j):
        return (i % nu) * nv + (j % nv) + 1
    for i in range(nu):
        for j in range(nv):
            a, b, c, d = vid(i, j), vid(i + 1, j), vid(i + 1, j + 1), vid(i, j + 1)
            faces.append((a, b, c))
            faces.append((a, c, d))
    with open(path, "w") as fh:
        for v in verts:
            fh.write("v %.6f %.6f %.6f\n" % v)
        for t in uvs:
            fh.write("vt %.6f %.6f\n" % t)
        for n in norms:
            fh.write("vn %.6f %.6f %.6f\n" % n)
        for f in faces:
            fh.write("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (f[0], f[0], f[0], f[1], f[1], f[1], f[2], f[2], f[2]))
    return len(faces)


@dataclass
class Scene:
    spheres: np.ndarray  # [n,4] float32: cx,cy,cz, radius MEMBER (ctor r*r, kernel.cu:287)
    lights: np.ndarray  # [m,7] float32
    texture: Sprite
    sky: Sprite
    aspect: np.float32 = REFERENCE_ASPECT
    sky_size: float = SKY_SIZE
    extent: float = 10.0
    name: str = ""
    meta: dict = field(default_factory=dict)
    # "next" primitives; the reference ships cube_count = plane_count = 0 (kernel.cu:1231)
    cubes: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), dtype=np.float32))   # [n,6] ctor args c1.xyz, c2.xyz
    planes: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), dtype=np.float32))  # [n,6] pos.xyz, normal.xyz
    mesh: "Mesh | None" = None   # triangle mesh + flat BVH

    @property
    def n_spheres(self) -> int:
        return int(self.spheres.shape[0])


def smooth_texture(width: int, height: int, seed: int) -> Sprite:
    """Low-frequency sinusoid texture quantised to bytes (values ~[0.2,0.9])."""
    rng = Lcg(seed)
    ph = [rng.next() / 32768.0 * 2 * math.pi for _ in range(6)]
    v, u = np.meshgrid(np.arange(height) / height, np.arange(width) / width, indexing="ij")
    chans = []
    for c in range(3):
        f = 0.55 + 0.20 * np.sin(2 * math.pi * ((c + 1) * u + v) + ph[c]) + 0.15 * np.cos(
            2 * math.pi * (u - (c + 2) * v) + ph[3 + c]
        )
        chans.append(np.clip(np.rint(f * 255), 0, 255).astype(np.uint8))
    return Sprite.from_bytes(np.stack(chans, axis=-1))


def checker_texture(width: int, height: int, cells: int = 16) -> Sprite:
    """High-contrast checker (worst case for texel flips)."""
    v, u = np.meshgrid(np.arange(height) * cells // height, np.arange(width) * cells // width, indexing="ij")
    on = ((u + v) & 1).astype(np.uint8)
    rgb = np.stack([40 + 200 * on, 220 - 180 * on, 60 + 120 * on], axis=-1).astype(np.uint8)
    return Sprite.from_bytes(rgb)


def _sphere_draws(n: int, seed: int):
    rng = Lcg(seed)
    pos = np.empty((n, 3), dtype=np.int64)
    rad = np.empty(n, dtype=np.int64)
    for i in range(n):  # draw order x,y,z,r per sphere (kernel.cu:1190)
        pos[i, 0] = rng.next() % 100
        pos[i, 1] = rng.next() % 100
        pos[i, 2] = rng.next() % 100
        rad[i] = rng.next() % 100
    return pos, rad


def reference_scene(n: int, seed: int = 1, texture: Sprite | None = None, sky: Sprite | None = None) -> Scene:
    """R(N, seed): the reference's own generator, kernel.cu:1189-1191.

    centre = (rand()%100)/10, ctor radius r = (rand()%100)/100, stored member = r*r.
    """
    pos, rad = _sphere_draws(n, seed)
    c = pos.astype(np.float32) / np.float32(10)
    r = rad.astype(np.float32) / np.float32(100)
    spheres = np.concatenate([c, (r * r)[:, None]], axis=1).astype(np.float32)
    return Scene(
        spheres=np.ascontiguousarray(spheres),
        lights=REFERENCE_LIGHTS.copy(),
        texture=texture or smooth_texture(512, 512, seed + 101),
        sky=sky or smooth_texture(1024, 512, seed + 202),
        extent=10.0,
        name=f"R({n},{seed})",
    )


def scaled_scene(n: int, seed: int, texture: Sprite | None = None, sky: Sprite | None = None) -> Scene:
    """S(N, seed): same radius law, cube edge E = 10*(N/64)^(1/3) (constant sphere density),
    lights scaled by E/10."""
    pos, rad = _sphere_draws(n, seed)
    extent = 10.0 * (n / 64.0) ** (1.0 / 3.0)
    k = np.float32(extent / 10.0)
    c = (pos.astype(np.float32) / np.float32(10)) * k
    r = rad.astype(np.float32) / np.float32(100)
    spheres = np.concatenate([c, (r * r)[:, None]], axis=1).astype(np.float32)
    lights = REFERENCE_LIGHTS.copy()
    lights[:, :4] *= k
    return Scene(
        spheres=np.ascontiguousarray(spheres),
        lights=lights,
        texture=texture or smooth_texture(512, 512, seed + 101),
        sky=sky or smooth_texture(1024, 512, seed + 202),
        extent=extent,
        name=f"S({n},{seed})",
    )


def with_cubes_and_plane(scene: Scene, n_cubes: int, seed: int, plane: bool = True, reference_formula: bool = False) -> Scene:
    """Adds axis-aligned cubes and (optionally) the reference's floor plane to a scene.

    reference_formula=True reproduces object::loadMesh (kernel.cu:1193-1199): pos = (rand()%100)/0.9 per axis,
    cube(pos, pos-2) - far outside the sphere cloud; otherwise 2x2x2 cubes are scattered inside the scene's extent.
    The plane is the reference's `plane({0,-4,0}, normalise({0,1,0}))` (kernel.cu:1187), lifted to y = 0.5 for the
    scattered variant so that it is visible from the orbit cameras.
    """
    rng = Lcg(seed)
    cubes = np.zeros((n_cubes, 6), dtype=np.float32)
    for i in range(n_cubes):
        if reference_formula:
            p = [np.float32(float(rng.next() % 100) / 0.9) for _ in range(3)]
            cubes[i, :3] = p
            cubes[i, 3:] = [np.float32(v - np.float32(2)) for v in p]
        else:
            p = [np.float32((rng.next() % 100) / 100.0 * scene.extent) for _ in range(3)]
            e = np.float32(0.4 + (rng.next() % 100) / 100.0)
            cubes[i, :3] = [np.float32(v + e) for v in p]   # c1 is the upper corner, as in the reference (c2 = c1 - 2)
            cubes[i, 3:] = [np.float32(v - e) for v in p]
    planes = np.zeros((1 if plane else 0, 6), dtype=np.float32)
    if plane:
        planes[0] = [0, -4 if reference_formula else 0.5, 0, 0, 1, 0]
    out = Scene(spheres=scene.spheres, lights=scene.lights, texture=scene.texture, sky=scene.sky, aspect=scene.aspect,
                sky_size=scene.sky_size, extent=scene.extent, name=scene.name + f"+{n_cubes}cubes" + ("+plane" if plane else ""),
                cubes=cubes, planes=planes)
    return out


def orbit_camera(scene: Scene, frame: int, n_frames: int = 240, pitch_deg: float = 15.0) -> Camera:
    """Camera on an orbit of radius 1.2*E about the cube centre, facing the centre.

    Per camera::rotateDir (kernel.cu:252-255) the view axis (0,0,1) maps to
    (cos p * sin yaw, -sin p, cos p * cos yaw): yaw 180 looks down -z and a positive pitch
    looks down.  yaw = 180 + theta, theta = 360*frame/n_frames.
    """
    e = scene.extent
    yaw = 180.0 + 360.0 * frame / n_frames
    yr, pr = math.radians(yaw), math.radians(pitch_deg)
    axis = (math.cos(pr) * math.sin(yr), -math.sin(pr), math.cos(pr) * math.cos(yr))
    centre = (e / 2, e / 2, e / 2)
    org = tuple(float(np.float32(centre[i] - 1.2 * e * axis[i])) for i in range(3))
    return Camera(org=org, yaw=float(np.float32(yaw)), pitch=float(np.float32(pitch_deg)))


def reference_camera() -> Camera:
    """kernel.cu:1695 + :261 - the reference's start-up camera."""
    return Camera()
