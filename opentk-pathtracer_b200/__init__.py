"""opentk-pathtracer_b200 — a B200 (sm_100a) CUDA path-tracing integrator behind the OpenTK-PathTracer host surface.

Scope: the one hot path of BoyBaykiller/OpenTK-PathTracer — res/shaders/PathTracing/compute.glsl and the
src/Render/PathTracer.cs dispatch — as libptb200.so (C ABI in include/ptb200.h), plus the host-side mirror of the
reference's interface for that path.  Import as `import ptb200` (shim at the repo root) or
`importlib.import_module("opentk-pathtracer_b200")`.
"""
from . import scene
from ._build import build as build_library
from ._lib import PtbError, lib_path, load as load_library
from .pathtracer import (FORMAT_RGB32F, FORMAT_RGBA8, FORMAT_RGBA32F, KERNEL_MEGA, KERNEL_NAIVE, PRECISION_EXACT, PRECISION_FAST, AtmosphericScatterer, BufferObject,
                         PathTracer, ScreenEffect)
from .scene import Camera, Cuboid, Material, Scene, Sphere, default_camera, load_default_scene, synthetic_scene

__all__ = ["scene", "build_library", "load_library", "lib_path", "PtbError", "PathTracer", "ScreenEffect", "AtmosphericScatterer", "BufferObject", "KERNEL_MEGA",
           "KERNEL_NAIVE", "PRECISION_EXACT", "PRECISION_FAST", "FORMAT_RGBA32F", "FORMAT_RGB32F", "FORMAT_RGBA8", "Camera", "Cuboid", "Material", "Scene", "Sphere", "default_camera", "load_default_scene",
           "synthetic_scene"]
