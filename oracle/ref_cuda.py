"""ctypes loader for oracle/_ref/libglsl_ref_cuda[_fast].so — the reference's compute.glsl compiled by nvcc and dispatched
in the reference's own launch shape (oracle/build_ref.py --cuda, oracle/ref_harness_pt_cuda.inc).  TEST / MEASUREMENT
INFRASTRUCTURE ONLY: the GL-compute proxy built from the reference's source.  Needs a GPU; the product never loads it."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_ref
from .oracle import Params, _fp


class CudaReference:
    def __init__(self, fast: bool = False, capacity=None):
        self.path = build_ref.cuda_lib(fast, capacity)
        if not os.path.exists(self.path):
            raise FileNotFoundError(f"{self.path} is not built (python oracle/build_ref.py --cuda{' --fast' if fast else ''})")
        L = C.CDLL(self.path)
        fp = C.POINTER(C.c_float)
        L.glref_cuda_pt_render.argtypes = [C.POINTER(Params), C.c_void_p, C.c_void_p, fp, C.c_int, fp, fp]
        L.glref_cuda_pt_render.restype = C.c_int
        L.glref_cuda_pt_capacity.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.glref_cuda_last_error.restype = C.c_char_p
        self._L = L

    def capacity(self):
        s, c = C.c_int(), C.c_int()
        self._L.glref_cuda_pt_capacity(C.byref(s), C.byref(c))
        return s.value, c.value

    def render(self, image: np.ndarray, basic_ubo: bytes, objects_ubo: bytes, env: np.ndarray, *, frame: int, frames: int = 1,
               spp: int, ray_depth: int, focal_length: float, aperture_diameter: float, n_spheres: float, n_cuboids: float,
               max_spheres: int = 256) -> float:
        """`frames` dispatches starting at `frame` over `image` (H x W x 4 float32, in place); returns ms per dispatch
        (CUDA events around the dispatches only)."""
        assert image.dtype == np.float32 and image.ndim == 3 and image.shape[2] == 4 and image.flags.c_contiguous
        assert env.dtype == np.float32 and env.ndim == 4 and env.shape[0] == 6 and env.flags.c_contiguous
        cap_s, cap_c = self.capacity()
        blob = bytes(objects_ubo)
        blob += b"\0" * max(0, 80 * cap_s + 96 * cap_c - len(blob))
        h, w = image.shape[:2]
        p = Params(w, h, frame, spp, ray_depth, focal_length, aperture_diameter, float(n_spheres), float(n_cuboids), max_spheres,
                   env.shape[1], 0, h, 0, w, 0, 1)
        ms = C.c_float()
        rc = self._L.glref_cuda_pt_render(C.byref(p), C.create_string_buffer(bytes(basic_ubo), len(basic_ubo)),
                                          C.create_string_buffer(blob, len(blob)), _fp(env), frames, _fp(image), C.byref(ms))
        if rc != 0:
            raise RuntimeError(f"glref_cuda_pt_render failed: {rc} {self._L.glref_cuda_last_error().decode()}")
        return float(ms.value)
