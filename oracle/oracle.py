"""ctypes loader for the CPU oracle (oracle/libpt_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs — never by the product.
Parity pin: oracle/ref.py (the reference's shaders compiled for the CPU) and tests/test_reference_pin.py; see oracle/glsl_model.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpt_oracle.so")
_SOURCES = ("pt_oracle.c", "atmosphere_oracle.c", "glsl_model.h", "Makefile")


class Params(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("frame", C.c_int), ("spp", C.c_int), ("ray_depth", C.c_int),
                ("focal_length", C.c_float), ("aperture_diameter", C.c_float),
                ("n_spheres", C.c_float), ("n_cuboids", C.c_float), ("max_spheres", C.c_int), ("env_size", C.c_int),
                ("y0", C.c_int), ("y1", C.c_int), ("x0", C.c_int), ("x1", C.c_int), ("n_threads", C.c_int), ("y_step", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("bounces", C.c_uint64), ("hits", C.c_uint64), ("rng_draws", C.c_uint64),
                ("nonfinite_pixels", C.c_uint64), ("depth_hist", C.c_uint64 * 64)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile if the .so is missing or older than its sources."""
    stale = force or not os.path.exists(_LIB_PATH)
    if not stale:
        t = os.path.getmtime(_LIB_PATH)
        stale = any(os.path.getmtime(os.path.join(_HERE, s)) > t for s in _SOURCES)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        L.pto_render.argtypes = [C.POINTER(Params), C.c_void_p, C.c_void_p, fp, fp, C.POINTER(Stats)]
        L.pto_render.restype = C.c_int
        L.pto_max_threads.restype = C.c_int
        L.pto_seed.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
        L.pto_seed.restype = C.c_uint32
        L.pto_pcg_stream.argtypes = [C.c_uint32, C.c_int, u32p, fp]
        L.pto_sincos.argtypes = [fp, C.c_int, fp, fp]
        L.pto_exp.argtypes = [fp, C.c_int, fp]
        L.pto_ray_sphere.argtypes = [fp, C.c_int, C.c_void_p, fp]
        L.pto_ray_cuboid.argtypes = [fp, C.c_int, C.c_void_p, fp]
        L.pto_ray_trace.argtypes = [fp, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_float, fp]
        L.pto_texture_cube.argtypes = [fp, C.c_int, fp, C.c_int, fp]
        L.pto_atmosphere.argtypes = [C.c_int, C.c_void_p, fp, C.c_float, C.c_int, C.c_int, fp, C.c_int]
        L.pto_atmosphere.restype = C.c_int
        u8p = C.POINTER(C.c_uint8)
        L.pto_tonemap.argtypes = [fp, C.c_int, u8p]
        L.pto_srgb8_to_linear.argtypes = [u8p, C.c_int, fp]
        L.pto_log.argtypes = [fp, C.c_int, fp]
        _lib = L
    return _lib


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def max_threads() -> int:
    return int(lib().pto_max_threads())


def render(image: np.ndarray, basic_ubo: bytes, objects_ubo: bytes, env: np.ndarray, *, frame: int, spp: int,
           ray_depth: int, focal_length: float, aperture_diameter: float, n_spheres: int, n_cuboids: int,
           max_spheres: int = 256, rows=None, cols=None, n_threads: int = 0, want_stats: bool = False, y_step: int = 1):
    """One dispatch over `image` (H x W x 4 float32, updated in place: running mean, compute.glsl:126-129)."""
    assert image.dtype == np.float32 and image.ndim == 3 and image.shape[2] == 4 and image.flags.c_contiguous
    assert env.dtype == np.float32 and env.ndim == 4 and env.shape[0] == 6 and env.shape[3] == 4 and env.flags.c_contiguous
    h, w = image.shape[:2]
    y0, y1 = rows if rows is not None else (0, h)
    x0, x1 = cols if cols is not None else (0, w)
    p = Params(w, h, frame, spp, ray_depth, focal_length, aperture_diameter, float(n_spheres), float(n_cuboids),
               max_spheres, env.shape[1], y0, y1, x0, x1, n_threads, y_step)
    st = Stats() if want_stats else None
    b0 = C.create_string_buffer(bytes(basic_ubo), len(basic_ubo))
    b1 = C.create_string_buffer(bytes(objects_ubo), len(objects_ubo))
    rc = lib().pto_render(C.byref(p), b0, b1, _fp(env), _fp(image), C.byref(st) if st is not None else None)
    if rc != 0:
        raise RuntimeError(f"pto_render failed: {rc}")
    if st is None:
        return None
    return dict(samples=st.samples, bounces=st.bounces, hits=st.hits, rng_draws=st.rng_draws,
                nonfinite_pixels=st.nonfinite_pixels, depth_hist=list(st.depth_hist))


def seed(x: int, y: int, frame: int) -> int:
    return int(lib().pto_seed(x, y, frame))


def pcg_stream(seed_value: int, n: int):
    h = np.zeros(n, dtype=np.uint32)
    f = np.zeros(n, dtype=np.float32)
    lib().pto_pcg_stream(seed_value, n, h.ctypes.data_as(C.POINTER(C.c_uint32)), _fp(f))
    return h, f


def sincos(x: np.ndarray):
    x = np.ascontiguousarray(x, dtype=np.float32)
    s, c = np.empty_like(x), np.empty_like(x)
    lib().pto_sincos(_fp(x), x.size, _fp(s), _fp(c))
    return s, c


def exp(x: np.ndarray):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().pto_exp(_fp(x), x.size, _fp(y))
    return y


def ray_sphere(rays: np.ndarray, sphere80: bytes):
    rays = np.ascontiguousarray(rays, dtype=np.float32)
    out = np.empty((rays.shape[0], 4), dtype=np.float32)
    lib().pto_ray_sphere(_fp(rays), rays.shape[0], C.create_string_buffer(sphere80, 80), _fp(out))
    return out


def ray_cuboid(rays: np.ndarray, cuboid96: bytes):
    rays = np.ascontiguousarray(rays, dtype=np.float32)
    out = np.empty((rays.shape[0], 4), dtype=np.float32)
    lib().pto_ray_cuboid(_fp(rays), rays.shape[0], C.create_string_buffer(cuboid96, 96), _fp(out))
    return out


def ray_trace(rays: np.ndarray, objects_ubo: bytes, max_spheres: int, n_spheres: int, n_cuboids: int):
    rays = np.ascontiguousarray(rays, dtype=np.float32)
    out = np.empty((rays.shape[0], 12), dtype=np.float32)
    lib().pto_ray_trace(_fp(rays), rays.shape[0], C.create_string_buffer(objects_ubo, len(objects_ubo)), max_spheres,
                        float(n_spheres), float(n_cuboids), _fp(out))
    return out


def texture_cube(env: np.ndarray, dirs: np.ndarray):
    env = np.ascontiguousarray(env, dtype=np.float32)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32)
    out = np.empty((dirs.shape[0], 3), dtype=np.float32)
    lib().pto_texture_cube(_fp(env), env.shape[1], _fp(dirs), dirs.shape[0], _fp(out))
    return out


def atmosphere(size: int, ubo: bytes, light_pos, light_intensity: float, i_steps: int, j_steps: int, n_threads: int = 0):
    out = np.empty((6, size, size, 4), dtype=np.float32)
    lp = np.ascontiguousarray(light_pos, dtype=np.float32)
    rc = lib().pto_atmosphere(size, C.create_string_buffer(ubo, len(ubo)), _fp(lp), light_intensity, i_steps, j_steps,
                              _fp(out), n_threads)
    if rc != 0:
        raise RuntimeError(f"pto_atmosphere failed: {rc}")
    return out


def tonemap(image: np.ndarray) -> np.ndarray:
    """ScreenEffect.Render(PathTracer.Result): ACES fit + linear->sRGB into RGBA8 (PostProcessing/fragment.glsl)."""
    a = np.ascontiguousarray(image, dtype=np.float32)
    out = np.empty(a.shape[:-1] + (4,), dtype=np.uint8)
    lib().pto_tonemap(_fp(a), int(a.size // 4), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def srgb8_to_linear(faces: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(faces, dtype=np.uint8)
    out = np.empty(a.shape, dtype=np.float32)
    lib().pto_srgb8_to_linear(a.ctypes.data_as(C.POINTER(C.c_uint8)), int(a.size // 4), _fp(out))
    return out


def log(x: np.ndarray):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().pto_log(_fp(x), x.size, _fp(y))
    return y
