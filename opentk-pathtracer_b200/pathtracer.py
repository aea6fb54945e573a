"""Host-side mirror of the reference's `PathTracer` class (src/Render/PathTracer.cs:9-141) over libptb200.so.

Same member names, argument meaning and call order as the C# class so the parity tests read like a port of
MainWindow's usage: properties push on set (PathTracer.cs:11-83), Render() dispatches once and bumps the frame
counter (:114-129), SetSize() re-allocates and resets (:131-135), ResetRenderer() only zeroes the counter
(:137-140), Samples = frames * SPP (:112).  `BufferObject` mirrors SubData on the two UBOs
(src/Render/Objects/BufferObject.cs:37-48).  There is no CPU fallback: construction fails if the CUDA library
cannot be loaded or no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import scene as _scene

FORMAT_RGBA32F, FORMAT_RGB32F, FORMAT_RGBA8 = 0, 1, 2      # ptb_read_result_format_async
KERNEL_MEGA = 0
KERNEL_NAIVE = 1
PRECISION_EXACT, PRECISION_FAST = 0, 1                      # ptb_set_precision


def _as_bytes(data) -> bytes:
    if isinstance(data, (bytes, bytearray, memoryview)):
        return bytes(data)
    return np.ascontiguousarray(data).tobytes()


class BufferObject:
    """One of the two uniform buffers the integrator reads (binding 0 = BasicDataUBO, 1 = GameObjectsUBO)."""

    def __init__(self, tracer: "PathTracer", binding: int, size: int):
        self._tracer, self.binding, self.Size = tracer, binding, size

    def SubData(self, offset: int, size: int, data) -> None:
        raw = _as_bytes(data)[:size].ljust(size, b"\0")
        buf = C.create_string_buffer(raw, len(raw))
        L = _lib.load()
        fn = L.ptb_basic_data_subdata if self.binding == 0 else L.ptb_game_objects_subdata
        _lib.check(fn(self._tracer._ctx, offset, size, buf))


class ScreenEffect:
    """Mirror of src/Render/ScreenEffect.cs for the post-process pass: Render(tracer) tone-maps the tracer's Result
    (ACES fit + linear->sRGB, PostProcessing/fragment.glsl) into an RGBA8 image, `Result` (H, W, 4) uint8."""

    def __init__(self):
        self.Result = None

    def Render(self, tracer: "PathTracer") -> np.ndarray:
        rows = _lib.check(tracer._L.ptb_local_rows(tracer._ctx))
        out = np.empty((rows, tracer.Width, 4), dtype=np.uint8)
        if rows:
            _lib.check(tracer._L.ptb_tonemap_rgba8(tracer._ctx, out.ctypes.data_as(C.c_void_p)))
        self.Result = out
        return out


class AtmosphericScatterer:
    """Mirror of src/Render/AtmosphericScatterer.cs for the environment producer: the four push-on-set properties
    (ISteps, JSteps, Time, LightIntensity — the last clamped at 0 like the C# setter, :51), the constructor defaults
    (:91-94), Render() and SetSize().  Render() runs the atmosphere kernel of `tracer` and binds the cubemap as its
    EnvironmentMap (MainWindow.cs:174-175,189 do those two steps separately); Result reads the faces back."""

    def __init__(self, tracer: "PathTracer", size: int):
        if size <= 0:
            raise ValueError("size must be positive")
        self._tracer, self.Size = tracer, int(size)
        self.Fast = False            # True: the live-regeneration kernel (ptb_generate_atmosphere_fast); parity runs keep False
        self.Time = 0.5
        self.ISteps = 50
        self.JSteps = 15
        self.LightIntensity = 15.0

    @property
    def LightIntensity(self) -> float:
        return self._lightIntensity

    @LightIntensity.setter
    def LightIntensity(self, value: float) -> None:
        self._lightIntensity = max(float(value), 0.0)

    @property
    def LightPos(self) -> np.ndarray:
        """The `lightPos` uniform the Time setter uploads (AtmosphericScatterer.cs:41)."""
        return _scene.atmosphere_light_pos(self.Time)

    def Render(self) -> None:
        self._tracer.GenerateAtmosphere(self.Size, int(self.ISteps), int(self.JSteps), float(self.Time), self._lightIntensity, fast=self.Fast)

    def SetSize(self, size: int) -> None:
        if size <= 0:
            raise ValueError("size must be positive")
        self.Size = int(size)

    @property
    def Result(self) -> np.ndarray:
        return self._tracer.ReadEnvironment()


class PathTracer:
    def __init__(self, environmentMap, width: int, height: int, rayDepth: int, spp: int, focalLength: float,
                 apertureDiamater: float, *, max_spheres: int = _scene.MAX_GAMEOBJECTS_SPHERES,
                 max_cuboids: int = _scene.MAX_GAMEOBJECTS_CUBOIDS, device: int = 0):
        self._L = _lib.load()
        ctx = C.c_void_p()
        _lib.check(self._L.ptb_create(C.byref(ctx), width, height, max_spheres, max_cuboids, device))
        self._ctx = ctx
        self.max_spheres, self.max_cuboids = max_spheres, max_cuboids
        self._numSpheres = self._numCuboids = 0
        self.BasicDataUBO = BufferObject(self, 0, _scene.BASIC_DATA_SIZE)
        self.GameObjectsUBO = BufferObject(self, 1, _scene.SPHERE_SIZE * max_spheres + _scene.CUBOID_SIZE * max_cuboids)
        self.RayDepth = rayDepth
        self.SPP = spp
        self.FocalLength = focalLength
        self.ApertureDiameter = apertureDiamater
        if environmentMap is not None:
            self.EnvironmentMap = environmentMap

    # ------------------------------------------------------------------ properties (PathTracer.cs:11-83)
    @property
    def NumSpheres(self) -> int:
        return self._numSpheres

    @NumSpheres.setter
    def NumSpheres(self, value: int) -> None:
        _lib.check(self._L.ptb_set_num_spheres(self._ctx, int(value)))
        self._numSpheres = int(value)

    @property
    def NumCuboids(self) -> int:
        return self._numCuboids

    @NumCuboids.setter
    def NumCuboids(self, value: int) -> None:
        _lib.check(self._L.ptb_set_num_cuboids(self._ctx, int(value)))
        self._numCuboids = int(value)

    @property
    def RayDepth(self) -> int:
        return self._rayDepth

    @RayDepth.setter
    def RayDepth(self, value: int) -> None:
        _lib.check(self._L.ptb_set_ray_depth(self._ctx, int(value)))
        self._rayDepth = int(value)

    @property
    def SPP(self) -> int:
        return self._spp

    @SPP.setter
    def SPP(self, value: int) -> None:
        _lib.check(self._L.ptb_set_spp(self._ctx, int(value)))
        self._spp = int(value)

    @property
    def FocalLength(self) -> float:
        return self._focalLength

    @FocalLength.setter
    def FocalLength(self, value: float) -> None:
        _lib.check(self._L.ptb_set_focal_length(self._ctx, float(value)))
        self._focalLength = float(value)

    @property
    def ApertureDiameter(self) -> float:
        return self._apertureDiameter

    @ApertureDiameter.setter
    def ApertureDiameter(self, value: float) -> None:
        _lib.check(self._L.ptb_set_aperture_diameter(self._ctx, float(value)))
        self._apertureDiameter = float(value)

    @property
    def EnvironmentMap(self):
        return self._env

    @EnvironmentMap.setter
    def EnvironmentMap(self, faces) -> None:
        """faces: float32 array (6, N, N, 4), faces +X,-X,+Y,-Y,+Z,-Z (the RGBA32F cubemap, PathTracer.cs:85,118)."""
        a = np.ascontiguousarray(faces, dtype=np.float32)
        if a.ndim != 4 or a.shape[0] != 6 or a.shape[1] != a.shape[2] or a.shape[3] != 4:
            raise ValueError("EnvironmentMap must be (6, N, N, 4) float32")
        _lib.check(self._L.ptb_set_environment_rgba32f(self._ctx, a.shape[1], a.ctypes.data_as(C.POINTER(C.c_float))))
        self._env = a

    def SetSkyBox(self, faces_rgba8) -> None:
        """EnvironmentMap = SkyBox: six sRGB8 faces (6, N, N, 4) uint8, decoded to linear on upload (Helper.cs:18-50, Gui.cs:84-86)."""
        a = np.ascontiguousarray(faces_rgba8, dtype=np.uint8)
        if a.ndim != 4 or a.shape[0] != 6 or a.shape[1] != a.shape[2] or a.shape[3] != 4:
            raise ValueError("SkyBox must be (6, N, N, 4) uint8")
        _lib.check(self._L.ptb_set_environment_srgb8(self._ctx, a.shape[1], a.ctypes.data_as(C.c_void_p)))
        self._env = None

    def GenerateAtmosphere(self, size: int = 256, iSteps: int = 50, jSteps: int = 15, time: float = 0.5,
                           lightIntensity: float = 15.0, fast: bool = False) -> None:
        """AtmosphericScatterer(size).Render() on the GPU, result bound as the EnvironmentMap (MainWindow.cs:174-175,189).
        fast=True: the live-regeneration kernel (tabulated secondary loop, ~1e-4 relative to the exact one)."""
        ubo = _scene.atmosphere_ubo_bytes()
        lp = np.ascontiguousarray(_scene.atmosphere_light_pos(time), dtype=np.float32)
        fn = self._L.ptb_generate_atmosphere_fast if fast else self._L.ptb_generate_atmosphere
        _lib.check(fn(self._ctx, size, C.create_string_buffer(ubo, len(ubo)), len(ubo),
                                                   lp.ctypes.data_as(C.POINTER(C.c_float)), float(lightIntensity), iSteps, jSteps))
        self._env = None

    def ReadEnvironment(self) -> np.ndarray:
        n = _lib.check(self._L.ptb_environment_size(self._ctx))
        out = np.empty((6, n, n, 4), dtype=np.float32)
        _lib.check(self._L.ptb_read_environment(self._ctx, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    # ------------------------------------------------------------------ PathTracer.cs:112-140
    @property
    def Samples(self) -> int:
        return _lib.check(self._L.ptb_samples(self._ctx))

    @property
    def Frame(self) -> int:
        return _lib.check(self._L.ptb_frame(self._ctx))

    def Render(self, frames: int = 1) -> None:
        _lib.check(self._L.ptb_render_frames(self._ctx, frames) if frames != 1 else self._L.ptb_render(self._ctx))

    def SetSize(self, width: int, height: int) -> None:
        _lib.check(self._L.ptb_set_size(self._ctx, width, height))

    def ResetRenderer(self) -> None:
        _lib.check(self._L.ptb_reset(self._ctx))

    @property
    def Width(self) -> int:
        return _lib.check(self._L.ptb_width(self._ctx))

    @property
    def Height(self) -> int:
        return _lib.check(self._L.ptb_height(self._ctx))

    @property
    def Result(self) -> np.ndarray:
        """The RGBA32F accumulation image (local rows x W x 4), copied to the host."""
        rows = _lib.check(self._L.ptb_local_rows(self._ctx))
        out = np.empty((rows, self.Width, 4), dtype=np.float32)
        if rows:
            _lib.check(self._L.ptb_read_result(self._ctx, out.ctypes.data_as(C.c_void_p)))
        return out

    # ------------------------------------------------------------------ extras beyond the C# class
    def LoadScene(self, sc: "_scene.Scene") -> None:
        """MainWindow.LoadScene's upload loop (MainWindow.cs:211-212,265-266): counts, then one SubData per object."""
        self.NumSpheres = 0
        self.NumCuboids = 0
        for o in sc.spheres:
            o.Upload(self.GameObjectsUBO)
            self.NumSpheres = self.NumSpheres + 1
        for o in sc.cuboids:
            o.Upload(self.GameObjectsUBO)
            self.NumCuboids = self.NumCuboids + 1

    def SetCamera(self, camera: "_scene.Camera", fov=_scene.FOV) -> None:
        """The BasicDataUBO writes of MainWindow.cs:131-132 and :278-279."""
        data = _scene.basic_data_bytes(camera, self.Width, self.Height, fov)
        self.BasicDataUBO.SubData(0, 64, data[0:64])
        self.BasicDataUBO.SubData(64, 64, data[64:128])
        self.BasicDataUBO.SubData(128, 16, data[128:144])

    def SetKernel(self, kernel: int) -> None:
        _lib.check(self._L.ptb_set_kernel(self._ctx, kernel))

    def SetOverlap(self, n: int) -> None:
        """Frames in flight (ptb_set_overlap): >= 2 pipelines consecutive Render() calls, <= 1 renders in place."""
        _lib.check(self._L.ptb_set_overlap(self._ctx, n))

    def SetBatch(self, frames: int) -> None:
        """Render(n) traces up to `frames` consecutive frames per megakernel launch (ptb_set_batch; 1 = off)."""
        _lib.check(self._L.ptb_set_batch(self._ctx, frames))

    def SetGridDivisor(self, d: int) -> None:
        """Each frame's persistent grid takes 1/d of the resident CTA slots (ptb_set_grid_divisor; experiment knob, default 1)."""
        _lib.check(self._L.ptb_set_grid_divisor(self._ctx, d))

    def SetTile(self, rank: int, world: int, stripe_rows: int = 8) -> None:
        _lib.check(self._L.ptb_set_tile(self._ctx, rank, world, stripe_rows))

    def SetFrame(self, frame: int) -> None:
        _lib.check(self._L.ptb_set_frame(self._ctx, frame))

    def WriteResult(self, image: np.ndarray) -> None:
        a = np.ascontiguousarray(image, dtype=np.float32)
        _lib.check(self._L.ptb_write_result(self._ctx, a.ctypes.data_as(C.c_void_p)))

    def ReadResultAsync(self, pinned_host_ptr: int, format: int = FORMAT_RGBA32F) -> None:
        """Enqueue a pipelined read-back of the image into pinned host memory (valid after Synchronize()).
        format: FORMAT_RGBA32F (16 B/pixel), FORMAT_RGB32F (12 B/pixel: the constant alpha 1.0 stays on the device) or
        FORMAT_RGBA8 (4 B/pixel: the tone-mapped display frame, ScreenEffect fused into the snapshot)."""
        _lib.check(self._L.ptb_read_result_format_async(self._ctx, int(format), C.c_void_p(pinned_host_ptr)))

    def ReadResultScatterAsync(self, pinned_full_frame_ptr: int, format: int = FORMAT_RGBA32F) -> None:
        """Multi-GPU read-back: this rank's stripes go straight into their rows of ONE full-frame host image (a pinned,
        usually shared mapping), over this GPU's own PCIe link (ptb_read_result_scatter_async)."""
        _lib.check(self._L.ptb_read_result_scatter_async(self._ctx, int(format), C.c_void_p(pinned_full_frame_ptr)))

    def RegisterGLTexture(self, texture: int) -> None:
        """CUDA-GL interop: `texture` is the Rgba32f TEXTURE_2D the host samples as PathTracer.Result (needs a current GL context)."""
        _lib.check(self._L.ptb_register_gl_texture(self._ctx, int(texture)))

    def PresentGL(self) -> None:
        _lib.check(self._L.ptb_present_gl(self._ctx))

    def Synchronize(self) -> None:
        _lib.check(self._L.ptb_synchronize(self._ctx))

    def LastRenderMs(self) -> float:
        return float(self._L.ptb_last_render_ms(self._ctx))

    @property
    def KernelLaunches(self) -> int:
        return _lib.check(self._L.ptb_kernel_launches(self._ctx))

    def SetKernelTiming(self, enabled: bool) -> None:
        _lib.check(self._L.ptb_set_kernel_timing(self._ctx, int(enabled)))

    def KernelTime(self) -> dict:
        """Summed device time of the megakernel launches since the last call (ptb_kernel_time): ms, frames, launches."""
        ms, fr, ln = C.c_double(), C.c_longlong(), C.c_longlong()
        _lib.check(self._L.ptb_kernel_time(self._ctx, C.byref(ms), C.byref(fr), C.byref(ln)))
        return dict(ms=ms.value, frames=fr.value, launches=ln.value)

    def SetPrecision(self, precision: int) -> None:
        """PRECISION_EXACT (default; bit-identical to the oracle) or PRECISION_FAST (MUFU + FMA build of the same kernel)."""
        _lib.check(self._L.ptb_set_precision(self._ctx, int(precision)))

    @property
    def Precision(self) -> int:
        return _lib.check(self._L.ptb_precision(self._ctx))

    def SetRayClassification(self, mode: int = 1, cells: int = 18, buckets: int = 16) -> None:
        """Ray-classification table for scenes of <= 64 primitives (ptb_set_ray_classification); mode 0 = plain fold."""
        _lib.check(self._L.ptb_set_ray_classification(self._ctx, int(mode), int(cells), int(buckets)))

    def SetLargeSceneMode(self, mode: int) -> None:
        """Scenes above the BVH threshold: 1 = uniform grid + DDA (default), 0 = binary BVH (ptb_set_large_scene_mode)."""
        _lib.check(self._L.ptb_set_large_scene_mode(self._ctx, int(mode)))

    def SetGridDensity(self, cells_per_primitive: float) -> None:
        _lib.check(self._L.ptb_set_grid_density(self._ctx, float(cells_per_primitive)))

    def SetBvhThreshold(self, primitives: int) -> None:
        _lib.check(self._L.ptb_set_bvh_threshold(self._ctx, int(primitives)))

    def SceneInfo(self, what: int) -> int:
        return _lib.check(self._L.ptb_scene_info(self._ctx, int(what)))

    @property
    def BvhNodes(self) -> int:
        """Nodes of the shared-memory BVH the current scene was packed into (0 = brute-force fold)."""
        return self.SceneInfo(0)

    def ResultDevicePtr(self):
        p, n = C.c_void_p(), C.c_size_t()
        _lib.check(self._L.ptb_result_device_ptr(self._ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    def SetStream(self, cuda_stream: int) -> None:
        _lib.check(self._L.ptb_set_stream(self._ctx, C.c_void_p(cuda_stream)))

    def SetStats(self, enabled: bool) -> None:
        _lib.check(self._L.ptb_set_stats(self._ctx, int(enabled)))

    def ReadStats(self) -> dict:
        out = (C.c_ulonglong * 3)()
        _lib.check(self._L.ptb_read_stats(self._ctx, out))
        return dict(samples=out[0], bounces=out[1], hits=out[2])

    def DebugEval(self, op: int, data: np.ndarray, n: int, out_floats: int) -> np.ndarray:
        a = np.ascontiguousarray(data, dtype=np.float32)
        out = np.empty(out_floats, dtype=np.float32)
        _lib.check(self._L.ptb_debug_eval(self._ctx, op, a.ctypes.data_as(C.POINTER(C.c_float)), n,
                                          out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def Dispose(self) -> None:
        if getattr(self, "_ctx", None):
            self._L.ptb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.Dispose()
        except Exception:
            pass
