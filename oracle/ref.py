"""ctypes loader for oracle/_ref/libglsl_ref.so — the reference's own GLSL shaders compiled for the CPU by
oracle/build_ref.py (through oracle/glsl_shim.hpp).  TEST INFRASTRUCTURE ONLY, same rule as oracle/oracle.py: tests/,
__graft_entry__ and bench.py's reference arm may load it; the product never does.

The library is built where /root/reference exists (this container; `__graft_entry__.build()` does it) and travels to the
GPU box as a prebuilt binary (oracle/_ref/ is git-ignored, not gpurun-ignored).  Nothing here reads /root/reference at
run time.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from .oracle import Params, _fp

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libglsl_ref.so")
MANIFEST_PATH = os.path.join(_HERE, "_ref", "manifest.json")

_lib = None


def use_library(path: str) -> None:
    """Point this module at another build of the compiled shaders (build_ref.py --capacity / --alt-model)."""
    global LIB_PATH, _lib
    LIB_PATH, _lib = path, None


def variant(path: str):
    """A second, independent instance of this module bound to another build of the compiled shaders."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(f"oracle.ref_variant_{abs(hash(path))}", os.path.abspath(__file__))
    m = importlib.util.module_from_spec(spec)
    m.__package__ = "oracle"
    spec.loader.exec_module(m)
    m.use_library(path)
    return m


def available() -> bool:
    return os.path.exists(LIB_PATH)


def manifest() -> dict:
    with open(MANIFEST_PATH) as f:
        return json.load(f)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f"{LIB_PATH} is not built (python oracle/build_ref.py, needs /root/reference)")
        L = C.CDLL(LIB_PATH)
        fp, u32p, u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
        L.glref_pt_render.argtypes = [C.POINTER(Params), C.c_void_p, C.c_void_p, fp, fp]
        L.glref_pt_render.restype = C.c_int
        L.glref_pt_capacity.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.glref_pt_pcg_stream.argtypes = [C.c_uint32, C.c_int, u32p, fp]
        L.glref_texture_cube.argtypes = [fp, C.c_int, fp, C.c_int, fp]
        L.glref_pt_load_scene.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.glref_pt_load_scene.restype = C.c_int
        L.glref_pt_ray_trace.argtypes = [fp, C.c_int, fp]
        L.glref_atmosphere.argtypes = [C.c_int, C.c_void_p, fp, C.c_float, C.c_int, C.c_int, fp, C.c_int]
        L.glref_atmosphere.restype = C.c_int
        L.glref_post.argtypes = [fp, C.c_int, C.c_int, u8p]
        L.glref_post.restype = C.c_int
        _lib = L
    return _lib


def capacity() -> tuple[int, int]:
    """(Spheres[], Cuboids[]) array lengths the shader declares (compute.glsl:69-70)."""
    s, c = C.c_int(), C.c_int()
    lib().glref_pt_capacity(C.byref(s), C.byref(c))
    return s.value, c.value


def render(image: np.ndarray, basic_ubo: bytes, objects_ubo: bytes, env: np.ndarray, *, frame: int, spp: int,
           ray_depth: int, focal_length: float, aperture_diameter: float, n_spheres: int, n_cuboids: int,
           max_spheres: int = 256, rows=None, cols=None, n_threads: int = 0, y_step: int = 1) -> None:
    """One dispatch of the compiled compute.glsl over `image` (H x W x 4 float32, read and written in place).
    Same signature as oracle.render so the two can be driven by the same test code."""
    assert image.dtype == np.float32 and image.ndim == 3 and image.shape[2] == 4 and image.flags.c_contiguous
    assert env.dtype == np.float32 and env.ndim == 4 and env.shape[0] == 6 and env.shape[3] == 4 and env.flags.c_contiguous
    cap_s, cap_c = capacity()
    if max_spheres != cap_s:
        raise ValueError(f"the shader declares Spheres[{cap_s}]; max_spheres={max_spheres} is a different block layout")
    blob = _padded_ubo(objects_ubo)
    h, w = image.shape[:2]
    y0, y1 = rows if rows is not None else (0, h)
    x0, x1 = cols if cols is not None else (0, w)
    p = Params(w, h, frame, spp, ray_depth, focal_length, aperture_diameter, float(n_spheres), float(n_cuboids),
               max_spheres, env.shape[1], y0, y1, x0, x1, n_threads, y_step)
    b0 = C.create_string_buffer(bytes(basic_ubo), len(basic_ubo))
    b1 = C.create_string_buffer(blob, len(blob))
    rc = lib().glref_pt_render(C.byref(p), b0, b1, _fp(env), _fp(image))
    if rc != 0:
        raise RuntimeError(f"glref_pt_render failed: {rc}")


def pcg_stream(seed_value: int, n: int):
    h = np.zeros(n, dtype=np.uint32)
    f = np.zeros(n, dtype=np.float32)
    lib().glref_pt_pcg_stream(seed_value, n, h.ctypes.data_as(C.POINTER(C.c_uint32)), _fp(f))
    return h, f


def _padded_ubo(objects_ubo: bytes) -> bytes:
    cap_s, cap_c = capacity()
    blob = bytes(objects_ubo)
    return blob + b"\0" * max(0, 80 * cap_s + 96 * cap_c - len(blob))


def texture_cube(env: np.ndarray, dirs: np.ndarray) -> np.ndarray:
    """texture(samplerCube, dir).rgb through the shim's own texture unit (n x 3)."""
    env = np.ascontiguousarray(env, dtype=np.float32)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32)
    out = np.empty((dirs.shape[0], 4), dtype=np.float32)
    lib().glref_texture_cube(_fp(env), env.shape[1], _fp(dirs), dirs.shape[0], _fp(out))
    return np.ascontiguousarray(out[:, :3])


def ray_trace(rays: np.ndarray, objects_ubo: bytes, max_spheres: int, n_spheres: float, n_cuboids: float) -> np.ndarray:
    """The shader's RayTrace() for n rays; same output format as oracle.ray_trace."""
    rays = np.ascontiguousarray(rays, dtype=np.float32)
    blob = _padded_ubo(objects_ubo)
    rc = lib().glref_pt_load_scene(C.create_string_buffer(blob, len(blob)), max_spheres, float(n_spheres), float(n_cuboids))
    if rc != 0:
        raise ValueError(f"glref_pt_load_scene failed: {rc}")
    out = np.empty((rays.shape[0], 12), dtype=np.float32)
    lib().glref_pt_ray_trace(_fp(rays), rays.shape[0], _fp(out))
    return out


def atmosphere(size: int, ubo: bytes, light_pos, light_intensity: float, i_steps: int, j_steps: int, n_threads: int = 0):
    out = np.zeros((6, size, size, 4), dtype=np.float32)
    lp = np.ascontiguousarray(light_pos, dtype=np.float32)
    rc = lib().glref_atmosphere(size, C.create_string_buffer(bytes(ubo), len(ubo)), _fp(lp), light_intensity, i_steps,
                                j_steps, _fp(out), n_threads)
    if rc != 0:
        raise RuntimeError(f"glref_atmosphere failed: {rc}")
    return out


def post(image: np.ndarray) -> np.ndarray:
    """PostProcessing/fragment.glsl over an H x W x 4 float32 image -> H x W x 4 uint8."""
    a = np.ascontiguousarray(image, dtype=np.float32)
    assert a.ndim == 3 and a.shape[2] == 4
    out = np.empty(a.shape, dtype=np.uint8)
    rc = lib().glref_post(_fp(a), a.shape[1], a.shape[0], out.ctypes.data_as(C.POINTER(C.c_uint8)))
    if rc != 0:
        raise RuntimeError(f"glref_post failed: {rc}")
    return out
