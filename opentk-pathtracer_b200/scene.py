"""Host-side byte producers: the scene / camera data the OpenTK front-end hands to the path tracer.

This mirrors the reference's C# data surface (it is host logic, numpy only, no GPU):
  Material.GetGPUFriendlyData     src/Material.cs:36-51      (4 x vec4 = 64 B)
  Sphere.GetGPUFriendlyData       src/GameObjects/Sphere.cs:23-31   (80 B, offset Instance*80)
  Cuboid.GetGPUFriendlyData       src/GameObjects/Cuboid.cs:21-35   (96 B, offset MAX_SPHERES*80 + Instance*96)
  MainWindow.LoadScene            src/MainWindow.cs:208-267  (48 spheres + 7 cuboids)
  Camera / BasicDataUBO writes    src/Camera.cs:16-30,79-82; src/MainWindow.cs:131-132,278-279
  CPU mouse picking               src/Ray.cs:15-19, Sphere.cs:34-50, Cuboid.cs:38-52, MainWindow.cs:302-322, Gui.cs:229-233
  Skybox image loading            src/Helper.cs:18-50, MainWindow.cs:177-187 (six PNG faces -> RGBA8, uploaded as Srgb8Alpha8)
OpenTK 3.3.2's Matrix4 helpers (LookAt, CreatePerspectiveFieldOfView, Inverted) are a NuGet dependency that is
not vendored in the reference; they are restated here in float32 from their published algorithms.  Parity with
the shader is defined at the UBO-byte boundary, so these only have to be self-consistent.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

f32 = np.float32

MAX_GAMEOBJECTS_SPHERES = 256  # MainWindow.cs:17
MAX_GAMEOBJECTS_CUBOIDS = 64
HOST_EPSILON = f32(0.005)      # MainWindow.cs:18 (wall thickness; NOT the shader's 0.001)
FOV = f32(103.0)
NEAR_FAR = (f32(0.005), f32(1000.0))  # MainWindow.cs:32

MATERIAL_SIZE = 16 * 4          # Material.cs:9
SPHERE_SIZE = 16 + MATERIAL_SIZE      # Sphere.cs:8
CUBOID_SIZE = 16 * 2 + MATERIAL_SIZE  # Cuboid.cs:8
BASIC_DATA_SIZE = 16 * 4 * 2 + 16     # MainWindow.cs:196


def _v3(x, y, z):
    return np.array([x, y, z], dtype=f32)


# ----------------------------------------------------------------------------- OpenTK math (float32)
def degrees_to_radians(deg) -> np.float32:
    """MathHelper.DegreesToRadians(float): degrees * ((float)Math.PI / 180f)."""
    return f32(deg) * (f32(math.pi) / f32(180.0))


def _normalize(v):
    v = v.astype(f32)
    s = f32(1.0) / f32(np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]))
    return (v * s).astype(f32)


def _cross(a, b):
    return _v3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def _dot(a, b):
    return f32(a[0] * b[0] + a[1] * b[1] + a[2] * b[2])


def look_at(eye, target, up) -> np.ndarray:
    """Matrix4.LookAt (row-vector convention; rows returned as a 4x4 float32 array)."""
    eye, target, up = (np.asarray(t, dtype=f32) for t in (eye, target, up))
    z = _normalize(eye - target)
    x = _normalize(_cross(up, z))
    y = _normalize(_cross(z, x))
    m = np.zeros((4, 4), dtype=f32)
    m[0] = (x[0], y[0], z[0], 0)
    m[1] = (x[1], y[1], z[1], 0)
    m[2] = (x[2], y[2], z[2], 0)
    m[3] = (-_dot(x, eye), -_dot(y, eye), -_dot(z, eye), 1)
    return m


def create_perspective_fov(fovy, aspect, z_near, z_far) -> np.ndarray:
    """Matrix4.CreatePerspectiveFieldOfView -> CreatePerspectiveOffCenter."""
    fovy, aspect, z_near, z_far = f32(fovy), f32(aspect), f32(z_near), f32(z_far)
    y_max = z_near * f32(math.tan(float(f32(0.5) * fovy)))
    y_min = -y_max
    x_min = y_min * aspect
    x_max = y_max * aspect
    left, right, bottom, top = x_min, x_max, y_min, y_max
    x = (f32(2.0) * z_near) / (right - left)
    y = (f32(2.0) * z_near) / (top - bottom)
    a = (right + left) / (right - left)
    b = (top + bottom) / (top - bottom)
    c = -(z_far + z_near) / (z_far - z_near)
    d = -(f32(2.0) * z_far * z_near) / (z_far - z_near)
    m = np.zeros((4, 4), dtype=f32)
    m[0, 0] = x
    m[1, 1] = y
    m[2] = (a, b, c, -1)
    m[3, 2] = d
    return m


def inverted(mat: np.ndarray) -> np.ndarray:
    """Matrix4.Inverted: Gauss-Jordan elimination with full pivoting, carried out in float32."""
    inv = np.array(mat, dtype=f32).copy()
    col_idx = [0] * 4
    row_idx = [0] * 4
    pivot_idx = [-1] * 4
    icol = irow = 0
    for i in range(4):
        max_pivot = f32(0.0)
        for j in range(4):
            if pivot_idx[j] != 0:
                for k in range(4):
                    if pivot_idx[k] == -1:
                        a = abs(inv[j, k])
                        if a > max_pivot:
                            max_pivot = a
                            irow, icol = j, k
                    elif pivot_idx[k] > 0:
                        return np.array(mat, dtype=f32)
        pivot_idx[icol] += 1
        if irow != icol:
            inv[[irow, icol]] = inv[[icol, irow]]
        row_idx[i], col_idx[i] = irow, icol
        pivot = inv[icol, icol]
        if pivot == 0:
            raise ValueError("Matrix is singular and cannot be inverted.")
        one_over = f32(1.0) / pivot
        inv[icol, icol] = f32(1.0)
        inv[icol] = (inv[icol] * one_over).astype(f32)
        for j in range(4):
            if j != icol:
                fct = inv[j, icol]
                inv[j, icol] = f32(0.0)
                inv[j] = (inv[j] - inv[icol] * fct).astype(f32)
    for j in range(3, -1, -1):
        ir, ic = row_idx[j], col_idx[j]
        if ir != ic:
            inv[:, [ir, ic]] = inv[:, [ic, ir]]
    return inv


def matrix_bytes(m: np.ndarray) -> bytes:
    """What BufferObject.SubData(..., Matrix4) uploads: Row0..Row3, which GLSL reads as columns 0..3."""
    return np.ascontiguousarray(m, dtype=f32).tobytes()


# ----------------------------------------------------------------------------- scene objects
@dataclass
class Material:
    """src/Material.cs — ctor clamps at :26-29 apply only through `new Material(...)`, not field writes."""
    Albedo: np.ndarray
    Emissiv: np.ndarray
    AbsorbanceColor: np.ndarray
    SpecularChance: float
    SpecularRoughness: float
    IOR: float
    RefractionChance: float
    RefractionRoughnes: float

    @staticmethod
    def new(albedo, emissiv, refractionColor, specularChance, specularRoughness, indexOfRefraction,
            refractionChance, refractionRoughnes) -> "Material":
        spec = f32(min(max(f32(specularChance), f32(0.0)), f32(1.0)))
        return Material(
            Albedo=np.asarray(albedo, dtype=f32), Emissiv=np.asarray(emissiv, dtype=f32),
            AbsorbanceColor=np.asarray(refractionColor, dtype=f32),
            SpecularChance=spec, SpecularRoughness=f32(specularRoughness),
            IOR=f32(max(f32(indexOfRefraction), f32(1.0))),
            RefractionChance=f32(min(max(f32(refractionChance), f32(0.0)), f32(1.0) - spec)),
            RefractionRoughnes=f32(refractionRoughnes))

    @staticmethod
    def Zero() -> "Material":  # Material.cs:8
        return Material.new(_v3(1, 1, 1), _v3(0, 0, 0), _v3(0, 0, 0), 0.0, 0.0, 1.0, 0.0, 0.0)

    def GetGPUFriendlyData(self) -> np.ndarray:
        d = np.zeros((4, 4), dtype=f32)
        d[0, :3] = self.Albedo;          d[0, 3] = self.SpecularChance
        d[1, :3] = self.Emissiv;         d[1, 3] = self.SpecularRoughness
        d[2, :3] = self.AbsorbanceColor; d[2, 3] = self.RefractionChance
        d[3, 0] = self.RefractionRoughnes
        d[3, 1] = self.IOR
        return d


@dataclass
class Sphere:
    Position: np.ndarray
    Radius: float
    Instance: int
    Material: Material
    max_spheres: int = MAX_GAMEOBJECTS_SPHERES

    @property
    def BufferOffset(self) -> int:  # Sphere.cs:20
        return 0 + self.Instance * SPHERE_SIZE

    def GetGPUFriendlyData(self) -> np.ndarray:
        d = np.zeros((5, 4), dtype=f32)
        d[0, :3] = self.Position
        d[0, 3] = self.Radius
        d[1:] = self.Material.GetGPUFriendlyData()
        return d

    def Upload(self, buffer) -> None:  # BaseSTD140Compatible.cs:12-16
        data = self.GetGPUFriendlyData()
        buffer.SubData(self.BufferOffset, data.nbytes, data)


@dataclass
class Cuboid:
    Position: np.ndarray
    Dimensions: np.ndarray
    Instance: int
    Material: Material
    max_spheres: int = MAX_GAMEOBJECTS_SPHERES

    @property
    def BufferOffset(self) -> int:  # Cuboid.cs:21
        return SPHERE_SIZE * self.max_spheres + self.Instance * CUBOID_SIZE

    @property
    def Min(self):
        return (self.Position - self.Dimensions * f32(0.5)).astype(f32)

    @property
    def Max(self):
        return (self.Position + self.Dimensions * f32(0.5)).astype(f32)

    def GetGPUFriendlyData(self) -> np.ndarray:
        d = np.zeros((6, 4), dtype=f32)
        d[0, :3] = self.Min
        d[1, :3] = self.Max
        d[2:] = self.Material.GetGPUFriendlyData()
        return d

    def Upload(self, buffer) -> None:
        data = self.GetGPUFriendlyData()
        buffer.SubData(self.BufferOffset, data.nbytes, data)


class HostBuffer:
    """A CPU stand-in for BufferObject (src/Render/Objects/BufferObject.cs:37-48): SubData writes bytes."""

    def __init__(self, size: int):
        self.Size = size
        self.data = bytearray(size)

    def SubData(self, offset: int, size: int, data) -> None:
        raw = np.ascontiguousarray(data).tobytes() if not isinstance(data, (bytes, bytearray)) else bytes(data)
        if offset < 0 or size < 0 or offset + size > self.Size:
            raise ValueError("SubData range outside the buffer")
        raw = raw[:size].ljust(size, b"\0")
        self.data[offset:offset + size] = raw

    def bytes(self) -> bytes:
        return bytes(self.data)


@dataclass
class Scene:
    spheres: list = field(default_factory=list)
    cuboids: list = field(default_factory=list)
    max_spheres: int = MAX_GAMEOBJECTS_SPHERES
    max_cuboids: int = MAX_GAMEOBJECTS_CUBOIDS

    @property
    def ubo_size(self) -> int:  # MainWindow.cs:200
        return SPHERE_SIZE * self.max_spheres + CUBOID_SIZE * self.max_cuboids

    def objects(self):
        return list(self.spheres) + list(self.cuboids)

    def ubo_bytes(self) -> bytes:
        buf = HostBuffer(self.ubo_size)
        for o in self.objects():
            o.Upload(buf)
        return buf.bytes()


def load_default_scene() -> Scene:
    """MainWindow.LoadScene (MainWindow.cs:208-267), float32 step by step as the C# evaluates it."""
    sc = Scene()
    width, height, depth = f32(40.0), f32(25.0), f32(25.0)
    eps = HOST_EPSILON
    balls = 6
    radius = f32(1.3)
    dimensions = _v3(width * f32(0.6), height, depth)
    for xi in range(balls):
        for yi in range(balls):
            x, y = f32(xi), f32(yi)
            pos = _v3(dimensions[0] / f32(balls) * x * f32(1.1) - dimensions[0] / f32(2),
                      (dimensions[1] / f32(balls)) * y - dimensions[1] / f32(2) + radius,
                      f32(-5))
            mat = Material.new(_v3(0.59, 0.59, 0.99), _v3(0, 0, 0), _v3(0, 0, 0), x / f32(balls - 1),
                               y / f32(balls - 1), 1.0, 0.0, 0.1)
            sc.spheres.append(Sphere(pos, radius, len(sc.spheres), mat))
    delta = (dimensions / f32(balls)).astype(f32)
    for xi in range(balls):
        x = f32(xi)
        m = Material.Zero()
        m.Albedo = _v3(0.9, 0.25, 0.25)
        m.SpecularChance = f32(0.02)
        m.IOR = f32(1.05)
        m.RefractionChance = f32(0.98)
        m.AbsorbanceColor = (_v3(1, 2, 3) * (x / f32(balls))).astype(f32)
        pos = _v3(-dimensions[0] / f32(2) + radius + delta[0] * x, f32(3.0), f32(-20.0))
        sc.spheres.append(Sphere(pos, radius, len(sc.spheres), m))
        m1 = Material.Zero()
        m1.SpecularChance = f32(0.02)
        m1.SpecularRoughness = x / f32(balls)
        m1.IOR = f32(1.1)
        m1.RefractionChance = f32(0.98)
        m1.RefractionRoughnes = x / f32(balls)
        m1.AbsorbanceColor = _v3(0, 0, 0)
        pos = _v3(-dimensions[0] / f32(2) + radius + delta[0] * x, f32(-6.0), f32(-20.0))
        sc.spheres.append(Sphere(pos, radius, len(sc.spheres), m1))

    def cub(pos, dims, mat):
        c = Cuboid(np.asarray(pos, dtype=f32), np.asarray(dims, dtype=f32), len(sc.cuboids), mat)
        sc.cuboids.append(c)
        return c

    z3 = _v3(0, 0, 0)
    down = cub(_v3(0.0, -height / f32(2.0), -10.0), _v3(width, eps, depth),
               Material.new(_v3(0.2, 0.04, 0.04), z3, z3, 0.0, 0.051, 1.0, 0.0, 0.0))
    dp, dd = down.Position, down.Dimensions
    cub(_v3(0.0, f32(18.495) - eps, -4.0), _v3(dd[0] * f32(0.3), eps, dd[2] * f32(0.3)),
        Material.new(_v3(0.04, 0.04, 0.04), (_v3(0.917, 0.945, 0.513) * f32(5.0)).astype(f32), z3, 0.0, 0.0, 1.0, 0.0, 0.0))
    cub(_v3(dp[0], dp[1] + height / f32(2), dp[2] + depth / f32(2) - f32(5.0)), _v3(width, height, eps),
        Material.new(_v3(0.37109375, 0.67578125, 0.3359375), z3, z3, 0.0, 0.0, 1.0, 0.0, 0.0))
    cub(_v3(dp[0], dp[1] + height / f32(2) + eps, dp[2] - depth / f32(2)), _v3(width, height - eps * f32(2), 0.3),
        Material.new(_v3(1, 1, 1), z3, _v3(0.01, 0.01, 0.01), 0.04, 0.0, 1.0, 0.954, 0.0))
    cub(_v3(dp[0] + width / f32(2), dp[1] + height / f32(2.0), dp[2]), _v3(eps, height, depth),
        Material.new(_v3(0.9453125, 0.75390625, 0.3046875), z3, z3, 1.0, 0.19, 1.0, 0.0, 0.0))
    cub(_v3(dp[0] - width / f32(2), dp[1] + height / f32(2.0), dp[2]), _v3(eps, height, depth),
        Material.new(_v3(0.074219, 0.25, 0.453125), z3, z3, 0.0, 0.0, 1.0, 0.0, 0.0))
    cub(_v3(-15.0, f32(-10.5) + eps, -15.0), _v3(3.0, 6.0, 3.0),
        Material.new(_v3(1, 1, 1), z3, z3, 0.0, 0.0, 1.0, 0.0, 0.0))
    return sc


def synthetic_scene(n_spheres: int = 1024, n_cuboids: int = 256, seed: int = 1234) -> Scene:
    """BASELINE config C3: the default room + random boxes and spheres with
    Material.GetRndMaterial-like materials (Material.cs:54-58), RefractionRoughness >= 0.05."""
    rng = np.random.default_rng(seed)
    base = load_default_scene()
    sc = Scene(max_spheres=max(n_spheres, 1), max_cuboids=max(n_cuboids, 1))
    lo, hi = np.array([-19, -11, -21], dtype=f32), np.array([19, 11, 1], dtype=f32)

    def rnd_material():
        emissive = rng.random() < 0.2
        albedo = rng.random(3).astype(f32)
        emis = rng.random(3).astype(f32) if emissive else np.zeros(3, dtype=f32)
        absorb = (rng.random(3).astype(f32) * f32(2.0)).astype(f32)
        m = Material.new(albedo, emis, absorb, f32(rng.random()) * f32(0.5), f32(rng.random()),
                         f32(rng.random()) + f32(1), f32(rng.random()) * f32(0.5), f32(rng.random()))
        m.RefractionRoughnes = f32(max(m.RefractionRoughnes, f32(0.05)))
        return m

    for c in base.cuboids[:min(7, n_cuboids)]:
        sc.cuboids.append(Cuboid(c.Position, c.Dimensions, len(sc.cuboids), c.Material, sc.max_spheres))
    while len(sc.cuboids) < n_cuboids:
        centre = (lo + rng.random(3).astype(f32) * (hi - lo)).astype(f32)
        dims = (f32(0.3) + rng.random(3).astype(f32) * f32(1.2)).astype(f32)
        sc.cuboids.append(Cuboid(centre, dims, len(sc.cuboids), rnd_material(), sc.max_spheres))
    while len(sc.spheres) < n_spheres:
        centre = (lo + rng.random(3).astype(f32) * (hi - lo)).astype(f32)
        radius = f32(0.2) + f32(rng.random()) * f32(0.6)
        sc.spheres.append(Sphere(centre, radius, len(sc.spheres), rnd_material(), sc.max_spheres))
    return sc


# ----------------------------------------------------------------------------- camera / BasicDataUBO
@dataclass
class Camera:
    """src/Camera.cs:16-30 (constructor) and :79-82 (GenerateMatrix)."""
    Position: np.ndarray
    Up: np.ndarray
    LookX: float = -90.0
    LookY: float = 0.0

    @property
    def ViewDir(self) -> np.ndarray:
        lx, ly = degrees_to_radians(self.LookX), degrees_to_radians(self.LookY)
        return _v3(f32(math.cos(lx)) * f32(math.cos(ly)), f32(math.sin(ly)), f32(math.sin(lx)) * f32(math.cos(ly)))

    @property
    def View(self) -> np.ndarray:
        pos = np.asarray(self.Position, dtype=f32)
        return look_at(pos, pos + self.ViewDir, self.Up)


def default_camera() -> Camera:  # MainWindow.cs:36
    return Camera(_v3(-17.14, 3.53, -8.62), _v3(0, 1, 0), -32.2, 0.8)


def basic_data_bytes(camera: Camera, width: int, height: int, fov=FOV) -> bytes:
    """BasicDataUBO (144 B): InvProjection @0 (MainWindow.cs:278-279), InvView @64, ViewPos @128 (:131-132)."""
    buf = HostBuffer(BASIC_DATA_SIZE)
    inv_proj = inverted(create_perspective_fov(degrees_to_radians(fov), f32(width) / f32(height), *NEAR_FAR))
    buf.SubData(0, 64, inv_proj)
    buf.SubData(64, 64, inverted(camera.View))
    buf.SubData(128, 16, np.append(np.asarray(camera.Position, dtype=f32), f32(0)))
    return buf.bytes()


# ----------------------------------------------------------------------------- CPU picking (host-side, float32 like the C#)
@dataclass
class Ray:
    """src/Ray.cs — the host-side ray used for mouse picking (not the shader's ray: IEEE division, Math.Max / Math.Min)."""
    Origin: np.ndarray
    Direction: np.ndarray

    def GetPoint(self, deltaTime) -> np.ndarray:  # Ray.cs:10-13
        return (np.asarray(self.Origin, f32) + np.asarray(self.Direction, f32) * f32(deltaTime)).astype(f32)

    @staticmethod
    def GetWorldSpaceRay(inverseProjection: np.ndarray, inverseView: np.ndarray, worldPosition, normalizedDeviceCoords) -> "Ray":
        """Ray.cs:15-19: row-vector products (`vector * matrix`), rayEye.zw = (-1, 0), normalised direction."""
        ndc = np.asarray(normalizedDeviceCoords, dtype=f32)
        eye = _row_times_matrix(np.array([ndc[0], ndc[1], f32(-1.0), f32(1.0)], dtype=f32), inverseProjection)
        eye[2], eye[3] = f32(-1.0), f32(0.0)
        d = _row_times_matrix(eye, inverseView)[:3]
        return Ray(np.asarray(worldPosition, dtype=f32).copy(), _normalize(d))


def _row_times_matrix(v: np.ndarray, m: np.ndarray) -> np.ndarray:
    """OpenTK `Vector4 * Matrix4`: result.X = v.X*M11 + v.Y*M21 + v.Z*M31 + v.W*M41, ... (float32, left to right)."""
    m = np.asarray(m, dtype=f32)
    out = np.zeros(4, dtype=f32)
    for c in range(4):
        out[c] = f32(f32(f32(v[0] * m[0, c]) + f32(v[1] * m[1, c])) + f32(v[2] * m[2, c])) + f32(v[3] * m[3, c])
    return out


def sphere_intersects_ray(s: Sphere, ray: Ray):
    """Sphere.IntersectsRay (Sphere.cs:34-50) -> (hit, t1, t2)."""
    v = (np.asarray(ray.Origin, f32) - np.asarray(s.Position, f32)).astype(f32)
    d = np.asarray(ray.Direction, f32)
    b = _dot(d, v)
    c = f32(_dot(v, v) - f32(f32(s.Radius) * f32(s.Radius)))
    disc = f32(f32(b * b) - c)
    if disc < 0:
        return False, f32(0), f32(0)
    sq = f32(np.sqrt(disc))
    return True, f32(-b - sq), f32(-b + sq)


def cuboid_intersects_ray(c: Cuboid, ray: Ray):
    """Cuboid.IntersectsRay (Cuboid.cs:38-52): slab test with IEEE division; Math.Max / Math.Min propagate NaN."""
    o, d = np.asarray(ray.Origin, f32), np.asarray(ray.Direction, f32)
    with np.errstate(all="ignore"):
        t0s = ((c.Min - o) / d).astype(f32)
        t1s = ((c.Max - o) / d).astype(f32)
        lo, hi = np.minimum(t0s, t1s), np.maximum(t0s, t1s)          # Vector3.ComponentMin / ComponentMax
        t1 = np.maximum(f32(np.finfo(f32).min), np.maximum(lo[0], np.maximum(lo[1], lo[2])))
        t2 = np.minimum(f32(np.finfo(f32).max), np.minimum(hi[0], np.minimum(hi[1], hi[2])))
    return bool(t1 <= t2), f32(t1), f32(t2)


def pick(scene: Scene, ray: Ray):
    """MainWindow.RayTrace (MainWindow.cs:302-318): the same order-dependent fold as the shader's (accept on
    hit && t2 > 0 && t1 < tMin, spheres before cuboids as LoadScene adds them) -> (object or None, t1, t2)."""
    t1 = t2 = f32(0)
    picked = None
    t_min = f32(np.finfo(f32).max)
    for o in scene.objects():
        hit, a, b = sphere_intersects_ray(o, ray) if isinstance(o, Sphere) else cuboid_intersects_ray(o, ray)
        if hit and b > 0 and a < t_min:
            t1, t2 = a, b
            t_min = b if a < 0 else a          # GetSmallestPositive, MainWindow.cs:319-322
            picked = o
    return picked, t1, t2


def pick_at_cursor(scene: Scene, camera: Camera, width: int, height: int, x: int, y_from_top: int, fov=FOV):
    """Gui.cs:229-233: window coordinates (origin top-left) -> NDC -> world ray -> RayTrace."""
    wy = height - y_from_top
    ndc = (np.array([x, wy], dtype=f32) / np.array([width, height], dtype=f32) * f32(2.0) - f32(1.0)).astype(f32)
    inv_proj = inverted(create_perspective_fov(degrees_to_radians(fov), f32(width) / f32(height), *NEAR_FAR))
    ray = Ray.GetWorldSpaceRay(inv_proj, inverted(camera.View), camera.Position, ndc)
    return pick(scene, ray) + (ray,)


# ----------------------------------------------------------------------------- skybox faces (the other EnvironmentMap)
SKYBOX_FACE_FILES = ("posx.png", "negx.png", "posy.png", "negy.png", "posz.png", "negz.png")   # MainWindow.cs:179-186: +X,-X,+Y,-Y,+Z,-Z


def load_cubemap_images(paths) -> np.ndarray:
    """Helper.ParallelLoadCubemapImages (Helper.cs:18-50): six image files -> (6, N, N, 4) uint8 RGBA faces, row 0 = the
    image's top row (`GetPixelRowSpan(0)` is uploaded as texel row 0), ready for PathTracer.SetSkyBox / ptb_set_environment_srgb8.
    Same checks, same messages: six paths, all present, square, equal sizes.  Decoding is `Image.Load<Rgba32>`: any PNG
    colour type becomes 8-bit RGBA (opaque alpha where the file has none); PNG is lossless, so Pillow yields the same bytes."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    paths = list(paths)
    if len(paths) != 6:
        raise ValueError("Number of images must be equal to six")
    if not all(os.path.exists(p) for p in paths):
        raise FileNotFoundError("At least on of the specified paths is invalid")
    from PIL import Image

    def load(path):
        with Image.open(path) as im:
            return np.asarray(im.convert("RGBA"), dtype=np.uint8)

    with ThreadPoolExecutor(max_workers=6) as pool:          # Parallel.For(0, 6, ...)
        images = list(pool.map(load, paths))
    if not all(im.shape[0] == im.shape[1] and im.shape[1] == images[0].shape[1] for im in images):
        raise ValueError("Individual cubemap textures must be squares and every texture must be of the same size")
    return np.ascontiguousarray(np.stack(images))


def skybox_paths(directory: str):
    """The six face paths of MainWindow.cs:179-186 under `directory`, resolved without regard to case (the host runs on a
    case-insensitive file system: the code says `posx.png`, the repository ships `posX.png`)."""
    import os
    present = {name.lower(): name for name in os.listdir(directory)}
    return [os.path.join(directory, present.get(f, f)) for f in SKYBOX_FACE_FILES]


# ----------------------------------------------------------------------------- atmosphere inputs
def atmosphere_ubo_bytes() -> bytes:
    """AtmosphericDataUBO (464 B): InvProjection + 6 InvView (AtmosphericScatterer.cs:75-89)."""
    inv_proj = inverted(create_perspective_fov(degrees_to_radians(90.0), 1.0, 0.1, 10.0))
    z = _v3(0, 0, 0)
    looks = [(_v3(1, 0, 0), _v3(0, -1, 0)), (_v3(-1, 0, 0), _v3(0, -1, 0)),
             (_v3(0, 1, 0), _v3(0, 0, 1)), (_v3(0, -1, 0), _v3(0, 0, -1)),
             (_v3(0, 0, 1), _v3(0, -1, 0)), (_v3(0, 0, -1), _v3(0, -1, 0))]
    out = matrix_bytes(inv_proj)
    for d, up in looks:
        out += matrix_bytes(inverted(look_at(z, z + d, up)))
    return out + b"\0" * 16


def atmosphere_light_pos(time: float = 0.5) -> np.ndarray:
    """AtmosphericScatterer.Time setter (AtmosphericScatterer.cs:41)."""
    rad = degrees_to_radians(f32(time) * f32(360.0))
    return (_v3(0.0, f32(math.sin(rad)), f32(math.cos(rad))) * f32(149600000e3)).astype(f32)


ATMOSPHERE_DEFAULTS = dict(size=256, iSteps=50, jSteps=15, time=0.5, lightIntensity=15.0)  # AtmosphericScatterer.cs:91-94, MainWindow.cs:174
PATHTRACER_DEFAULTS = dict(rayDepth=13, spp=1, focalLength=20.0, apertureDiameter=0.14)     # MainWindow.cs:189
