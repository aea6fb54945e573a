/*
 * ptb200.h — C ABI of libptb200.so: a B200 (sm_100a) CUDA replacement for the path-tracing pass of
 * BoyBaykiller/OpenTK-PathTracer, i.e. res/shaders/PathTracing/compute.glsl plus the dispatch in
 * src/Render/PathTracer.cs.  The reference has no FFI: the seam is the C# class `PathTracer` and the GL
 * binding points it consumes.  Each entry point below names the reference member it replaces
 * (paths relative to OpenTK-PathTracer/).  INTEGRATION.md shows the [DllImport] shim.
 *
 * Conventions: every call returns 0 on success or a negative PTB_E_* code; ptb_last_error() returns a
 * thread-local message.  Nothing throws or aborts across the ABI.  A context is bound to one CUDA device and
 * is not re-entrant (the reference drives everything from the single GameWindow thread, Program.cs:13).
 * All pointers are host pointers unless the name says `device`.
 */
#ifndef PTB200_H
#define PTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTB_OK 0
#define PTB_E_INVALID -1   /* bad argument / range */
#define PTB_E_CUDA -2      /* a CUDA runtime call failed */
#define PTB_E_STATE -3     /* call order (e.g. render before an environment map was set) */
#define PTB_E_NOMEM -4

/* Kernel variants.  PTB_KERNEL_MEGA is the product; PTB_KERNEL_NAIVE is the labelled "GL-compute proxy"
 * (one thread per pixel, 8x8 groups, scene read from the raw UBO bytes — how the GLSL dispatch is organised,
 * PathTracer.cs:121 / compute.glsl:8), kept only as a baseline to time beside the product. */
#define PTB_KERNEL_MEGA 0
#define PTB_KERNEL_NAIVE 1

typedef struct ptb_ctx ptb_ctx;

const char* ptb_last_error(void);
int ptb_version(void);

/* new PathTracer(env, width, height, ...) — PathTracer.cs:96-110; UBO capacities MainWindow.cs:17,195-201.
 * device = CUDA ordinal.  The result image is zero-initialised (GL leaves it undefined, Texture.cs:184). */
int ptb_create(ptb_ctx** out, int width, int height, int max_spheres, int max_cuboids, int device);
void ptb_destroy(ptb_ctx* ctx);

/* PathTracer.SetSize — PathTracer.cs:131-135 (re-allocates Result, frame counter = 0). */
int ptb_set_size(ptb_ctx* ctx, int width, int height);
/* PathTracer.ResetRenderer — PathTracer.cs:137-140 (frame counter = 0; image not cleared). */
int ptb_reset(ptb_ctx* ctx);

/* Property setters RayDepth / SPP / FocalLength / ApertureDiameter — PathTracer.cs:37-83
 * (uniforms rayDepth, SPP, focalLength, apertureDiameter, compute.glsl:90-94). */
int ptb_set_ray_depth(ptb_ctx* ctx, int ray_depth);
int ptb_set_spp(ptb_ctx* ctx, int spp);
int ptb_set_focal_length(ptb_ctx* ctx, float focal_length);
int ptb_set_aperture_diameter(ptb_ctx* ctx, float aperture_diameter);
/* NumSpheres / NumCuboids — PathTracer.cs:11-34 (uniform vec2 uboGameObjectsSize, compute.glsl:88). */
int ptb_set_num_spheres(ptb_ctx* ctx, int n);
int ptb_set_num_cuboids(ptb_ctx* ctx, int n);

/* BufferObject.SubData on UBO binding 0 (BasicDataUBO, 144 B: InvProjection @0, InvView @64, ViewPos @128)
 * — BufferObject.cs:37-48 as called from MainWindow.cs:131-132,279. */
int ptb_basic_data_subdata(ptb_ctx* ctx, int offset, int size, const void* data);
/* BufferObject.SubData on UBO binding 1 (GameObjectsUBO: Sphere[max_spheres] stride 80 @0, Cuboid[max_cuboids]
 * stride 96 @ max_spheres*80) — BaseSTD140Compatible.cs:12-16, Sphere.cs:20, Cuboid.cs:21. */
int ptb_game_objects_subdata(ptb_ctx* ctx, int offset, int size, const void* data);

/* PathTracer.EnvironmentMap = <RGBA32F cubemap> — PathTracer.cs:85,118.  six_faces = 6*face_size^2*4 floats,
 * faces +X,-X,+Y,-Y,+Z,-Z, row-major with row index = t.  Sampled LINEAR + seamless (compute.glsl:177). */
int ptb_set_environment_rgba32f(ptb_ctx* ctx, int face_size, const float* six_faces);
/* PathTracer.EnvironmentMap = SkyBox — the Srgb8Alpha8 cubemap built from six PNG faces (Helper.cs:18-50,
 * MainWindow.cs:177-187, Gui.cs:84-86).  six_faces = 6*face_size^2 RGBA8 texels, same face / row order; decoded
 * sRGB -> linear before filtering, as the texture unit does. */
int ptb_set_environment_srgb8(ptb_ctx* ctx, int face_size, const unsigned char* six_faces);
/* AtmosphericScatterer(size).Render() into the environment map, on the GPU — AtmosphericScatterer.cs:63-113,
 * AtmosphericScattering/compute.glsl.  ubo = AtmosphericDataUBO bytes (InvProjection + 6 InvView, 448 B used);
 * light_pos / light_intensity / i_steps / j_steps are that shader's uniforms. */
int ptb_generate_atmosphere(ptb_ctx* ctx, int face_size, const void* ubo, int ubo_size, const float* light_pos,
                            float light_intensity, int i_steps, int j_steps);
/* The same pass for live regeneration (the GUI re-runs it on every slider tick, up to 2048^2 faces; Gui.cs:93-143): the same
 * loops with special-function exp / sqrt / 1/x and fused multiply-adds — a third of the instructions; agrees with
 * ptb_generate_atmosphere (which stays bit-exact with the shader and is this one's checker) to ~1e-5 relative.  Both reuse the
 * cubemap buffers when the size is unchanged.  Not for parity runs. */
int ptb_generate_atmosphere_fast(ptb_ctx* ctx, int face_size, const void* ubo, int ubo_size, const float* light_pos,
                                 float light_intensity, int i_steps, int j_steps);
/* Copies the current environment map (unpadded, 6*face_size^2*4 floats) back to the host. */
int ptb_read_environment(ptb_ctx* ctx, float* six_faces);
int ptb_environment_size(ptb_ctx* ctx);

/* PathTracer.Render() — PathTracer.cs:114-129: one dispatch, then thisRenderNumFrame++.  Asynchronous on the
 * context's stream. */
int ptb_render(ptb_ctx* ctx);
/* n consecutive Render() calls (n dispatches, same results), enqueued back to back. */
int ptb_render_frames(ptb_ctx* ctx, int n);
/* PathTracer.Samples — PathTracer.cs:112 (thisRenderNumFrame * SPP). */
int ptb_samples(ptb_ctx* ctx);
int ptb_frame(ptb_ctx* ctx);
int ptb_set_frame(ptb_ctx* ctx, int frame); /* checkpoint/resume of a progressive render */

/* PathTracer.Result readback (the reference hands the GL texture to ScreenEffect, MainWindow.cs:51). */
int ptb_read_result(ptb_ctx* ctx, float* rgba32f);            /* synchronous: W*H*4 floats, row 0 = y 0 */
/* Pipelined read-back: snapshots the image on the render stream and copies the snapshot to (pinned) host memory on a
 * second stream, so the PCIe transfer of frame f overlaps Render() of frame f+1.  Up to two reads in flight; the data is
 * valid after ptb_synchronize(). */
int ptb_read_result_async(ptb_ctx* ctx, float* pinned_rgba32f);
/* The same pipelined read-back with the pixel format a host would ask glGetTexImage for (Texture.cs wraps Rgba32f; the
 * alpha the shader stores is the constant 1.0, compute.glsl:129, so shipping it over PCIe buys nothing):
 *   PTB_FORMAT_RGBA32F  16 B/pixel, identical to ptb_read_result_async;
 *   PTB_FORMAT_RGB32F   12 B/pixel, the three colour floats bit for bit, packed (W*H*3 floats);
 *   PTB_FORMAT_RGBA8     4 B/pixel, the display frame: ScreenEffect's tone-map pass (as ptb_tonemap_rgba8) fused into the snapshot.
 * The snapshot is a kernel on the render stream; the copy runs on the copy stream.  Valid after ptb_synchronize(). */
#define PTB_FORMAT_RGBA32F 0
#define PTB_FORMAT_RGB32F 1
#define PTB_FORMAT_RGBA8 2
int ptb_read_result_format_async(ptb_ctx* ctx, int format, void* pinned_dst);
/* Multi-GPU read-back (after ptb_set_tile): this rank's stripes go straight into THEIR ROWS of one full-frame host image
 * (`pinned_full_frame` = row 0 of the W x H frame in `format`; typically one shared mapping that every rank registered
 * with cudaHostRegister), one strided D2H per rank over that GPU's own PCIe link — the host-side frame is assembled
 * without any device-side exchange.  With world == 1 it equals ptb_read_result_format_async. */
int ptb_read_result_scatter_async(ptb_ctx* ctx, int format, void* pinned_full_frame);
int ptb_write_result(ptb_ctx* ctx, const float* rgba32f);     /* restore an accumulation image */
/* ScreenEffect.Render(PathTracer.Result) — ScreenEffect.cs:29-37 with PostProcessing/fragment.glsl: ACES fit + linear->sRGB
 * into an RGBA8 image (W*H*4 bytes, row 0 = y 0), the step right after the path tracer (MainWindow.cs:51). */
int ptb_tonemap_rgba8(ptb_ctx* ctx, unsigned char* rgba8);
int ptb_tonemap_device(ptb_ctx* ctx, void* rgba8_device);
int ptb_synchronize(ptb_ctx* ctx);

/* PathTracer.Result as the GL texture it is in the reference (PathTracer.cs:86,97-99: Rgba32f TEXTURE_2D, sampled by
 * ScreenEffect.Render, MainWindow.cs:51) through CUDA-GL interop: register the texture once (and again after SetSize
 * re-allocates it) from the thread that owns the GL context; ptb_present_gl copies the accumulation image into the texture on
 * the device, stream-ordered after the frames rendered so far — no host round trip.  Without a current GL context on the
 * context's GPU registration fails with PTB_E_STATE. */
int ptb_register_gl_texture(ptb_ctx* ctx, unsigned int gl_texture);
int ptb_present_gl(ptb_ctx* ctx);
int ptb_unregister_gl_texture(ptb_ctx* ctx);

/* Device-side access for hosts that own CUDA memory / streams (PyTorch, CUDA-GL interop). */
int ptb_result_device_ptr(ptb_ctx* ctx, void** device_ptr, size_t* bytes);
int ptb_set_stream(ptb_ctx* ctx, void* cuda_stream);
int ptb_width(ptb_ctx* ctx);
int ptb_height(ptb_ctx* ctx);

/* Pixel-tile partition for one-process-per-GPU rendering: this context renders only the rows y with
 * (y / stripe_rows) % world == rank, stored compactly (local row order) in its result image; seeds use global
 * pixel coordinates (compute.glsl:104-106) so the union over ranks is bit-identical to a single-GPU render. */
int ptb_set_tile(ptb_ctx* ctx, int rank, int world, int stripe_rows);
int ptb_local_rows(ptb_ctx* ctx);
/* Scatter rank-major gathered stripe buffers (world x max_local_rows x W x 4, device) into a full row-major
 * image (H x W x 4, device) — the de-interleave after the per-frame NCCL gather. */
int ptb_deinterleave_device(ptb_ctx* ctx, const void* gathered_device, void* full_device);
int ptb_max_local_rows(ptb_ctx* ctx);

/* Fused multi-GPU exchange (the alternative to an NCCL gather + de-interleave): every rank's blend kernel also stores its
 * blended pixels into rank 0's row-major full image through a CUDA-IPC peer mapping (NVLink stores), with system-scope
 * arrival / release flags for flow control.  Call order: ptb_set_tile, ptb_exchange_init(slots) on every rank;
 * ptb_exchange_handle on rank 0 -> ship the 64 bytes to the other ranks -> ptb_exchange_attach there; then per frame
 * ptb_render on every rank, and on rank 0 ptb_exchange_acquire (stream-ordered wait for all ranks, returns the frame's
 * device buffer), consumer work on the context stream, ptb_exchange_release.  Waits time out after 4 s (ptb_exchange_status). */
int ptb_exchange_init(ptb_ctx* ctx, int slots);
/* The same with the pixel format of the slots: PTB_FORMAT_RGBA32F (what ptb_exchange_init uses) or PTB_FORMAT_RGB32F — packed
 * colour floats, bit for bit; the constant alpha 1.0 (compute.glsl:129) is not shipped, which takes a quarter off rank 0's
 * NVLink ingress, the resource that bounds the exchange once the frame is split over many GPUs. */
int ptb_exchange_init_format(ptb_ctx* ctx, int slots, int format);
/* Rotating roots (rotate != 0): frame number q of the exchange is assembled on rank q % world instead of always on rank 0 —
 * one gather per frame as before, but no single GPU's NVLink ingress carries every frame (at 8 GPUs rank 0 would take in 7/8
 * of each frame: 22 MB per 1080p frame, the bound of the whole exchange once a frame takes ~25 us).  Every rank then owns a
 * block of `slots` images and maps all the others': ptb_exchange_init_roots on every rank, ptb_exchange_handle on every rank,
 * all-gather the handles, ptb_exchange_attach_peer(rank, handle) for each; every rank acquires / releases the frames it is
 * the root of (ptb_exchange_root tells which rank holds a given frame). */
int ptb_exchange_init_roots(ptb_ctx* ctx, int slots, int format, int rotate);
int ptb_exchange_attach_peer(ptb_ctx* ctx, int peer_rank, const void* handle64);
int ptb_exchange_root(ptb_ctx* ctx, long long frame_seq);   /* frame_seq < 0: the last frame rendered */
int ptb_exchange_pending(ptb_ctx* ctx);                      /* frames rendered that this rank is the root of and has not acquired yet */
int ptb_exchange_handle(ptb_ctx* ctx, void* handle64);
int ptb_exchange_attach(ptb_ctx* ctx, const void* handle64);
int ptb_exchange_acquire(ptb_ctx* ctx, void** full_device);
int ptb_exchange_release(ptb_ctx* ctx);
int ptb_exchange_status(ptb_ctx* ctx);

/* Kernel selection and introspection. */
int ptb_set_kernel(ptb_ctx* ctx, int kernel);
/* Frame pipelining.  n >= 2 (default 2): consecutive Render() calls are traced on n alternating streams into per-frame
 * scratch images and folded into the accumulation image by a stream-ordered blend kernel (same arithmetic as
 * compute.glsl:126-129), so the tail of frame f overlaps the start of frame f+1.  n <= 1: one stream, in-place blend. */
int ptb_set_overlap(ptb_ctx* ctx, int n);
/* Experiment knob (default 1 = every resident CTA slot): each frame's persistent grid takes 1/d of the slots (never fewer
 * than one CTA per SM), so that with n >= d frames in flight d frames are co-resident and one frame's drain runs beside
 * another's bulk instead of beside an empty machine.  Results do not depend on it. */
int ptb_set_grid_divisor(ptb_ctx* ctx, int d);
/* Frame batching (default 16; 1 = off).  ptb_render_frames(n) with n >= 2 traces up to `frames` consecutive frames per
 * megakernel launch: the work counter runs over all their pixels, frame-major, each frame's estimate goes to its own scratch
 * image, and ONE blend kernel folds the batch into the accumulation image frame by frame in registers — the same arithmetic
 * in the same order as n single-frame calls, so the image is bit-identical — but lanes move from the last pixels of one
 * frame straight into the next and only the last frame of a batch pays the drain of the longest paths (an 8-GPU share of a
 * 1080p frame: 67 -> 47 us/frame).  A plain ptb_render() is never batched.  Applies to the megakernel with overlap >= 2,
 * width <= 4096, statistics off; with the fused exchange every frame of a batch still lands in its own slot on rank 0, so a
 * batch is capped at the number of exchange slots.  Scratch memory: 2 x frames x image size, allocated at the first batch. */
int ptb_set_batch(ptb_ctx* ctx, int frames);
/* Arithmetic of the megakernel (protocols P1 / P2).
 *   PTB_PRECISION_EXACT (default): the evaluation model of DESIGN.md §2 — IEEE fp32, no contraction, correctly rounded 1/x and
 *     sqrt, polynomial sin/cos/exp — bit-identical to the CPU oracle and to the reference's shaders compiled for the CPU.
 *   PTB_PRECISION_FAST: the same kernel source compiled with fused multiply-adds and the GPU's special-function unit
 *     (MUFU rcp / rsq / sin / cos / ex2, ~1-2 ulp, denormals flushed) — what a GL driver's compiler may do with compute.glsl
 *     (GLSL 4.50 §4.7.1).  Same RNG stream and control flow; results differ from the exact build in the last bits of most
 *     samples and, where a discrete decision flips, in individual paths: held to per-channel MSE < 1e-6 against the exact
 *     build at matched seeds after 1024 frames (the north star's tolerance; tests/test_parity_gpu.py).
 * The proxy kernel, the statistics counters and the debug probes always run the exact arithmetic. */
#define PTB_PRECISION_EXACT 0
#define PTB_PRECISION_FAST 1
int ptb_set_precision(ptb_ctx* ctx, int precision);
int ptb_precision(ptb_ctx* ctx);
/* Ray classification for scenes of 4..64 primitives (default on): rays are classed by origin cell x direction bucket and a
 * table gives each class the set of primitives any of its rays can hit; RayTrace() then tests only those, in index order.
 * cells = grid cells along the longest scene axis (default 18; fewer when the table would exceed 64 MiB), buckets = direction
 * buckets per cube-face axis (default 16): 32.5 MB for the default scene.
 * The table is rebuilt on the GPU when geometry changes (not on material edits).  Results do not depend on it. */
int ptb_set_ray_classification(ptb_ctx* ctx, int mode, int cells, int buckets);
/* Large scenes (>= the BVH threshold): 1 (default) = uniform grid in shared memory walked by a 3-D DDA, 0 = binary BVH.  The
 * grid falls back to the BVH when its lists do not fit 16-bit offsets.  Results do not depend on it. */
int ptb_set_large_scene_mode(ptb_ctx* ctx, int mode);
int ptb_set_grid_density(ptb_ctx* ctx, float cells_per_primitive);   /* grid resolution: target cells per primitive (default 3; at most 32 per axis) */
/* Scenes with at least this many primitives are traced through the shared-memory BVH, smaller ones by the brute-force fold
 * (default 96).  Results do not depend on it (the hierarchy only removes primitives that fail the exact test). */
int ptb_set_bvh_threshold(ptb_ctx* ctx, int primitives);
/* Device time of the megakernel launches themselves, in whatever mode is running (pipelined, batched, tiled): with timing
 * enabled every launch records when its first CTA starts and when its last CTA ends (%globaltimer, two atomics per CTA);
 * ptb_kernel_time() waits for the outstanding launches, returns the summed duration in ms, the frames and the launches they
 * covered, and restarts the count.  Consecutive launches of a pipelined render overlap (the next grid's CTAs move in as the
 * previous grid's longest paths drain), so the sum can exceed the wall time of the region by that overlap. */
int ptb_set_kernel_timing(ptb_ctx* ctx, int enabled);
int ptb_kernel_time(ptb_ctx* ctx, double* ms_total, long long* frames, long long* launches);
int ptb_kernel_launches(ptb_ctx* ctx);          /* CUDA kernels launched by this context so far */
/* Introspection of the packed scene / launch shape (syncs the scene first): what the last LoadScene turned into. */
#define PTB_INFO_BVH_NODES 0      /* nodes of the shared-memory BVH (0: brute-force fold) */
#define PTB_INFO_ALWAYS_TESTED 1  /* primitives outside the hierarchy (scene-sized or non-finite), tested for every ray */
#define PTB_INFO_STAGED_BYTES 2   /* bytes of the scene block each CTA stages into shared memory */
#define PTB_INFO_GRID_CTAS 3      /* CTAs of the persistent grid of the last launch */
#define PTB_INFO_FOLD 4           /* how RayTrace() runs: 0 brute force, 1 BVH, 2 ray-classification table, 3 uniform grid */
#define PTB_INFO_RCT_KBYTES 5     /* size of the ray-classification table in KiB (0: none) */
#define PTB_INFO_GRID_CELLS 6     /* cells of the shared-memory grid (0: none) */
#define PTB_INFO_GRID_ITEMS 7     /* primitive references in the grid's cell lists */
int ptb_scene_info(ptb_ctx* ctx, int what);
float ptb_last_render_ms(ptb_ctx* ctx);         /* cudaEvent time of the last ptb_render[_frames] call (syncs) */
/* Path statistics of the next renders: counters[0]=samples, [1]=RayTrace calls, [2]=hits (device atomics; slow). */
int ptb_set_stats(ptb_ctx* ctx, int enabled);
int ptb_read_stats(ptb_ctx* ctx, unsigned long long* counters3);

/* Unit-level probes used by the parity tests: evaluate device functions on arrays (host in, host out).
 * op: 0 sincos (in n, out 2n)  1 exp (n -> n)  2 pcg stream (in: 1 seed as uint32 bits, out n floats)
 *     3 texture(samplerCube) (in 3n dirs, out 3n)  4 RayTrace fold over the packed scene (in 6n rays, out 12n)
 *     5 min/max/rcp/sqrt probe (in 2n, out 4n)  6 RayTrace fold over the raw UBO bytes (proxy view; as 4)
 *     8 log (n -> n)  9 RayTrace fold through the BVH (scenes of >= 96 primitives; as 4; out[12i+3] = nodes visited)
 *     10 RayTrace fold through the ray-classification table (as 4; out[12i+3] = candidates left, 65 = full mask)
 *     11 RayTrace fold through the uniform grid (large scenes; as 4; out[12i+3] = cells visited)
 *     7 group-cooperative fold of the frame tail: rays are processed k = in[6n] at a time per warp (in 6n+1, out 12n) */
int ptb_debug_eval(ptb_ctx* ctx, int op, const float* in, int n, float* out);

#ifdef __cplusplus
}
#endif
#endif
