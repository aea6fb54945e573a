"""ctypes binding of libptb200.so — exactly the symbols include/ptb200.h declares.  No fallback: if the CUDA
library is missing or fails to load, importing code gets an exception, never a CPU path."""
from __future__ import annotations

import ctypes as C
import os

from . import _build

_lib = None

# name -> (restype, argtypes); mirrors include/ptb200.h
_P = C.c_void_p
_FP = C.POINTER(C.c_float)
SIGNATURES = {
    "ptb_last_error": (C.c_char_p, []),
    "ptb_version": (C.c_int, []),
    "ptb_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ptb_destroy": (None, [_P]),
    "ptb_set_size": (C.c_int, [_P, C.c_int, C.c_int]),
    "ptb_reset": (C.c_int, [_P]),
    "ptb_set_ray_depth": (C.c_int, [_P, C.c_int]),
    "ptb_set_spp": (C.c_int, [_P, C.c_int]),
    "ptb_set_focal_length": (C.c_int, [_P, C.c_float]),
    "ptb_set_aperture_diameter": (C.c_int, [_P, C.c_float]),
    "ptb_set_num_spheres": (C.c_int, [_P, C.c_int]),
    "ptb_set_num_cuboids": (C.c_int, [_P, C.c_int]),
    "ptb_basic_data_subdata": (C.c_int, [_P, C.c_int, C.c_int, C.c_void_p]),
    "ptb_game_objects_subdata": (C.c_int, [_P, C.c_int, C.c_int, C.c_void_p]),
    "ptb_set_environment_rgba32f": (C.c_int, [_P, C.c_int, _FP]),
    "ptb_set_environment_srgb8": (C.c_int, [_P, C.c_int, C.c_void_p]),
    "ptb_tonemap_rgba8": (C.c_int, [_P, C.c_void_p]),
    "ptb_tonemap_device": (C.c_int, [_P, C.c_void_p]),
    "ptb_generate_atmosphere": (C.c_int, [_P, C.c_int, C.c_void_p, C.c_int, _FP, C.c_float, C.c_int, C.c_int]),
    "ptb_generate_atmosphere_fast": (C.c_int, [_P, C.c_int, C.c_void_p, C.c_int, _FP, C.c_float, C.c_int, C.c_int]),
    "ptb_read_environment": (C.c_int, [_P, _FP]),
    "ptb_environment_size": (C.c_int, [_P]),
    "ptb_render": (C.c_int, [_P]),
    "ptb_render_frames": (C.c_int, [_P, C.c_int]),
    "ptb_samples": (C.c_int, [_P]),
    "ptb_frame": (C.c_int, [_P]),
    "ptb_set_frame": (C.c_int, [_P, C.c_int]),
    "ptb_read_result": (C.c_int, [_P, C.c_void_p]),
    "ptb_read_result_async": (C.c_int, [_P, C.c_void_p]),
    "ptb_set_grid_divisor": (C.c_int, [_P, C.c_int]),
    "ptb_set_batch": (C.c_int, [_P, C.c_int]),
    "ptb_read_result_format_async": (C.c_int, [_P, C.c_int, C.c_void_p]),
    "ptb_read_result_scatter_async": (C.c_int, [_P, C.c_int, C.c_void_p]),
    "ptb_write_result": (C.c_int, [_P, C.c_void_p]),
    "ptb_register_gl_texture": (C.c_int, [_P, C.c_uint]),
    "ptb_present_gl": (C.c_int, [_P]),
    "ptb_unregister_gl_texture": (C.c_int, [_P]),
    "ptb_synchronize": (C.c_int, [_P]),
    "ptb_result_device_ptr": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "ptb_set_stream": (C.c_int, [_P, C.c_void_p]),
    "ptb_width": (C.c_int, [_P]),
    "ptb_height": (C.c_int, [_P]),
    "ptb_set_tile": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "ptb_local_rows": (C.c_int, [_P]),
    "ptb_max_local_rows": (C.c_int, [_P]),
    "ptb_deinterleave_device": (C.c_int, [_P, C.c_void_p, C.c_void_p]),
    "ptb_exchange_init": (C.c_int, [_P, C.c_int]),
    "ptb_exchange_init_format": (C.c_int, [_P, C.c_int, C.c_int]),
    "ptb_exchange_init_roots": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "ptb_exchange_attach_peer": (C.c_int, [_P, C.c_int, C.c_void_p]),
    "ptb_exchange_root": (C.c_int, [_P, C.c_longlong]),
    "ptb_exchange_pending": (C.c_int, [_P]),
    "ptb_exchange_handle": (C.c_int, [_P, C.c_void_p]),
    "ptb_exchange_attach": (C.c_int, [_P, C.c_void_p]),
    "ptb_exchange_acquire": (C.c_int, [_P, C.POINTER(C.c_void_p)]),
    "ptb_exchange_release": (C.c_int, [_P]),
    "ptb_exchange_status": (C.c_int, [_P]),
    "ptb_set_kernel": (C.c_int, [_P, C.c_int]),
    "ptb_set_overlap": (C.c_int, [_P, C.c_int]),
    "ptb_kernel_launches": (C.c_int, [_P]),
    "ptb_scene_info": (C.c_int, [_P, C.c_int]),
    "ptb_set_bvh_threshold": (C.c_int, [_P, C.c_int]),
    "ptb_set_large_scene_mode": (C.c_int, [_P, C.c_int]),
    "ptb_set_grid_density": (C.c_int, [_P, C.c_float]),
    "ptb_set_kernel_timing": (C.c_int, [_P, C.c_int]),
    "ptb_kernel_time": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "ptb_set_precision": (C.c_int, [_P, C.c_int]),
    "ptb_precision": (C.c_int, [_P]),
    "ptb_set_ray_classification": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "ptb_last_render_ms": (C.c_float, [_P]),
    "ptb_set_stats": (C.c_int, [_P, C.c_int]),
    "ptb_read_stats": (C.c_int, [_P, C.POINTER(C.c_ulonglong)]),
    "ptb_debug_eval": (C.c_int, [_P, C.c_int, _FP, C.c_int, _FP]),
}


def lib_path() -> str:
    return _build.LIB


def load():
    """dlopen libptb200.so (building it first if the sources are newer).  Raises if that is impossible."""
    global _lib
    if _lib is None:
        path = _build.build() if (_build.is_stale() and _can_build()) else _build.LIB
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing and cannot be built here; run __graft_entry__.build()")
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)   # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _can_build() -> bool:
    try:
        _build.nvcc()
        return True
    except RuntimeError:
        return False


class PtbError(RuntimeError):
    pass


def check(rc: int) -> int:
    if rc < 0:
        raise PtbError(f"libptb200 error {rc}: {load().ptb_last_error().decode(errors='replace')}")
    return rc
