#!/usr/bin/env python
"""bench.py — Msamples/s of the path-tracing pass on the default demo scene at 1920x1080 (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one PathTracer.Render() of the whole frame (one dispatch in the reference, PathTracer.cs:114-129): 1920*1080
pixels x SPP 1 = 2 073 600 samples, rayDepth 13, focal 20, aperture 0.14, default camera, 256^2 atmosphere cubemap.
At N > 1 the same frame is cut into interleaved 8-row stripes (one process per GPU), and every step ends with the one
exchange the path has: a gather of the stripe buffers to rank 0 + de-interleave.  Fixed total work => "strong" scaling.

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events on the launching stream, inputs resident in HBM, L2
flushed before every timed step; `e2e` goes through the same public API with host buffers: the per-frame camera UBO
upload the C# host does (MainWindow.cs:131-132) and a read-back of every frame into pinned host memory as RGB32F (the
colour floats bit for bit; the constant alpha 1.0 is not shipped) — at N > 1 each rank writes its stripes straight into one
full-frame image in host memory shared by all ranks, over its own PCIe link.  `gl_proxy` (N = 1) is the reference's own
compute.glsl compiled by nvcc and dispatched in the reference's launch shape, run in a subprocess after the timed regions.
`--impl reference` times the reference's own compute shader on the host CPUs: compute.glsl compiled by g++ from
/root/reference through oracle/build_ref.py (oracle/_ref/libglsl_ref.so, kind "reference"), all host threads; where that
binary is absent it falls back to the oracle's C restatement (kind "port").  The reference's C# + OpenGL host cannot run
here (no .NET, no GL).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
RAY_DEPTH, SPP, FOCAL, APERTURE = 13, 1, 20.0, 0.14
WORKLOAD = "default demo scene (48 spheres + 7 cuboids), 1920x1080, SPP 1 per frame, rayDepth 13, 256^2 atmosphere env (BASELINE configs[1])"
STRIPE_ROWS = 8
EXCHANGE_SLOTS = int(os.environ.get("PTB_SLOTS", "4"))      # full-image buffers on rank 0: how far the ranks may drift apart
L2_FLUSH_BYTES = 160 << 20      # larger than the 126 MB L2


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def issue_roofline(ncu, kern_ms, world):
    """The bound that actually binds: warp-instructions issued per second against 148 SMs x 4 schedulers x SM clock.
    Instruction count per launch comes from the committed ncu capture (full frame); duration is measured live."""
    inst = ncu.get("inst_executed_per_launch")
    if not inst:
        return None
    achieved = inst / world / (kern_ms * 1e-3) / 1e12
    peak = 148 * 4 * 1.965e9 / 1e12
    return {"achieved": achieved, "peak": peak, "unit": "T warp-inst/s", "frac": achieved / peak,
            "note": "peak = 148 SMs x 4 issue slots x 1.965 GHz; instructions per launch from profiles/ncu_summary.json"}


def ncu_summary():
    """Per-launch DRAM traffic of the megakernel from the committed `ncu --set full` capture (profiles/), if any."""
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return {}


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.01)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "n": len(self.samples)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ================================================================================================ reference arm
def host_threads():
    """Threads the CPU arm may use: the cores this process is allowed on.  torchrun exports OMP_NUM_THREADS=1 to its workers,
    which would silently turn the reference arm into a single-core run at N > 1 — the count is passed explicitly instead."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference():
    """(module, kind, description): the compiled reference shaders when oracle/_ref was built, else the oracle port."""
    from oracle import oracle as O, ref as R
    if R.available():
        return R, "reference", "the reference's compute.glsl compiled for the CPU (g++, oracle/build_ref.py -> oracle/_ref/libglsl_ref.so); its C# + OpenGL host needs .NET + GL 4.5, absent here"
    return O, "port", "CPU oracle (C restatement of compute.glsl); oracle/_ref not built on this machine"


def run_reference(args, rank, world):
    """The reference's shader on the host CPUs (oracle/_ref, kind "reference"; else the oracle, kind "port"), all threads,
    a bounded sample per step."""
    if rank != 0:
        return
    import ptb200
    from oracle import oracle as O_
    O, kind, kind_text = cpu_reference()
    sc = ptb200.scene
    threads = host_threads()
    env = O.atmosphere(256, sc.atmosphere_ubo_bytes(), sc.atmosphere_light_pos(0.5), 15.0, 50, 15, threads)
    scene, cam = sc.load_default_scene(), sc.default_camera()
    basic, ubo = sc.basic_data_bytes(cam, W, H), scene.ubo_bytes()
    img = np.zeros((H, W, 4), np.float32)
    kw = dict(spp=SPP, ray_depth=RAY_DEPTH, focal_length=FOCAL, aperture_diameter=APERTURE, n_spheres=48, n_cuboids=7, n_threads=threads)
    # calibrate on one full frame, then bound the per-step sample so warmup + steps stay under ~150 s
    t = time.perf_counter()
    O.render(img, basic, ubo, env, frame=0, **kw)
    full_s = time.perf_counter() - t
    budget = 150.0 / max(1, args.steps + args.warmup)
    y_step = max(1, min(64, int(np.ceil(full_s / budget))))          # every y_step-th row: spread over sky, spheres and floor
    rows_per_step = len(range(0, H, y_step))
    px_per_step = rows_per_step * W

    def step(frame):
        O.render(img, basic, ubo, env, frame=frame, y_step=y_step, **kw)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    dt = time.perf_counter() - t0
    value = px_per_step * SPP * args.steps / dt / 1e6
    sample = f"every {y_step}-th row of the 1920x1080 frame per step ({rows_per_step} rows, {px_per_step} samples/step)" if y_step > 1 else "the full 1920x1080 frame per step"
    line = {"impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "reference_kind": kind_text},
            "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ================================================================================================ our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import ptb200
    from importlib import import_module
    D = import_module("opentk-pathtracer_b200.distributed")

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sc = ptb200.scene
    scene, cam = sc.load_default_scene(), sc.default_camera()
    pt = ptb200.PathTracer(None, W, H, RAY_DEPTH, SPP, FOCAL, APERTURE, device=local_rank)
    pt.SetStream(torch.cuda.current_stream(dev).cuda_stream)
    # frames in flight: 2 on one GPU (a third starves the read-back snapshot of SM slots and costs e2e), 3 when the frame is
    # split over several GPUs (small per-GPU tiles are dominated by the per-frame tail; measured at N = 8: 22.1 -> 25.7 Gsamples/s)
    frames_in_flight = int(os.environ.get("PTB_OVERLAP", "3" if world > 1 else "2"))
    pt.SetOverlap(frames_in_flight)
    # opt-in experiment knobs (defaults leave everything as measured in round 1): PTB_BATCH = frames per megakernel launch
    # (ptb_set_batch; the device-timed loops then submit that many steps per call), PTB_GRID_DIV = ptb_set_grid_divisor
    batch = max(1, int(os.environ.get("PTB_BATCH", "1")))
    grid_div = max(1, int(os.environ.get("PTB_GRID_DIV", "1")))
    if batch > 1:
        pt.SetBatch(batch)
    if grid_div > 1:
        pt.SetGridDivisor(grid_div)
    pt.GenerateAtmosphere(256, 50, 15, 0.5, 15.0)      # the default EnvironmentMap, produced on the GPU (MainWindow.cs:174-175)
    pt.LoadScene(scene)
    pt.SetCamera(cam)
    # N > 1: the fused exchange (blend kernel stores straight into rank 0's image over NVLink); PTB_EXCHANGE=nccl selects the
    # NCCL gather + de-interleave path instead
    fused = os.environ.get("PTB_EXCHANGE", "fused") != "nccl" and world > 1
    tiled = None
    if world > 1:
        if fused:
            # CUDA IPC needs peer access between the ranks' devices; if any rank cannot set it up, everybody uses NCCL
            try:
                tiled = D.TiledPathTracer(pt, rank, world, STRIPE_ROWS, device=dev, fused=True, slots=EXCHANGE_SLOTS)
                ok = torch.ones(1, device=dev)
            except Exception as exc:      # noqa: BLE001
                print(f"rank {rank}: fused exchange unavailable ({exc}); falling back to the NCCL gather", file=sys.stderr, flush=True)
                ok = torch.zeros(1, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() < 1:
                fused, tiled = False, None
        if tiled is None:
            tiled = D.TiledPathTracer(pt, rank, world, STRIPE_ROWS, device=dev, fused=False)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    inv_view = sc.matrix_bytes(sc.inverted(cam.View))
    view_pos = np.append(np.asarray(cam.Position, np.float32), np.float32(0)).tobytes()
    rows_local = pt.Result.shape[0]

    # e2e read-back format: RGB32F — the colour floats bit for bit; the alpha the shader stores is the constant 1.0
    # (compute.glsl:129) and stays on the device.  N = 1: compact pipelined read into pinned memory.  N > 1: every rank writes
    # its stripes straight into ONE full-frame image in pinned host memory shared by all ranks (each over its own PCIe link);
    # if that mapping cannot be set up on every rank, the legacy path copies rank 0's assembled RGBA32F image instead.
    shared = None
    if world > 1 and os.environ.get("PTB_E2E", "scatter") == "scatter":
        try:
            shared = D.SharedHostFrame(W * H * 12, 2, rank, world)
        except Exception as exc:      # noqa: BLE001  (collective: raised on every rank or on none)
            if rank == 0:
                print(f"shared host frame unavailable ({exc}); e2e falls back to rank 0's copy of the assembled image", file=sys.stderr, flush=True)
            shared = None
    rgb_e2e = world == 1 or shared is not None
    host_bufs = [torch.empty((H if rank == 0 else 1, W, 3 if world == 1 else 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    side = torch.cuda.Stream(device=dev)
    copy_done = [torch.cuda.Event(), torch.cuda.Event()]
    snap = [torch.empty((H, W, 4), dtype=torch.float32, device=dev) for _ in range(2)] if (rank == 0 and world > 1) else None
    state = {"i": 0}

    def step_device(n=1):
        if tiled is None:
            pt.Render(n)
        elif n > 1 and fused:
            tiled.step_batch(n)    # one trace launch for n frames; every frame still reaches rank 0 through its own slot
        else:
            for _ in range(n):
                tiled.step()       # render f; finish gather f-1 (it overlapped this render); start gather f

    def finish_device():
        if tiled is not None:
            tiled.flush()

    def step_e2e():
        # what the C# host does every frame: camera UBO writes (host memory -> the library), Render(), and here the
        # accumulation image read back into pinned host memory (the reference hands it to the display pass instead).
        # The read-back is pipelined (device snapshot + copy on a second stream), so it overlaps the next Render().
        i = state["i"] = state["i"] + 1
        pt.BasicDataUBO.SubData(64, 64, inv_view)
        pt.BasicDataUBO.SubData(128, 16, view_pos)
        if tiled is None:
            pt.Render()
            pt.ReadResultAsync(host_bufs[i & 1].data_ptr(), ptb200.FORMAT_RGB32F)
        elif shared is not None:
            tiled.step()           # the device-side exchange keeps running exactly as in the device-timed loop
            pt.ReadResultScatterAsync(shared.ptr(i), ptb200.FORMAT_RGB32F)
        else:
            if fused:
                # rank 0 owns the slot between acquire and release: snapshot it (HBM->HBM) there, copy to the host on a side stream
                k = i & 1

                def consume(full, k=k):
                    cur = torch.cuda.current_stream(dev)
                    cur.wait_event(copy_done[k])             # the D2H that last read snap[k] is done
                    snap[k].copy_(full, non_blocking=True)
                    side.wait_stream(cur)
                    with torch.cuda.stream(side):
                        host_bufs[k].copy_(snap[k], non_blocking=True)
                        copy_done[k].record(side)
                tiled.step_fused(consumer=consume if rank == 0 else None)
            else:
                full = tiled.step()
                if rank == 0 and full is not None:
                    k = i & 1
                    cur = torch.cuda.current_stream(dev)
                    side.wait_stream(cur)
                    with torch.cuda.stream(side):
                        host_bufs[k].copy_(full, non_blocking=True)
                        copy_done[k].record(side)
                    cur.wait_event(copy_done[k])      # (cheap) keeps `full[k]` from being rewritten before its copy was issued

    def finish_e2e():
        if tiled is None:
            pt.Synchronize()
        elif shared is not None:
            tiled.flush()
            pt.Synchronize()       # render stream + this rank's copy stream; the closing barrier covers the other ranks
        else:
            full = tiled.flush()
            if rank == 0 and not fused:
                host_bufs[0].copy_(full, non_blocking=True)
            side.synchronize()
            torch.cuda.current_stream(dev).synchronize()

    def samples_per_step_():
        return W * H * SPP

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    flush_stream = torch.cuda.Stream(device=dev)

    def timed(fn, steps, flush_l2, finisher=None):
        """K steps enqueued back to back on the launching stream (the library pipelines consecutive frames), one CUDA-event
        bracket around the whole region.  L2 flush: a 160 MiB memset (> the 126 MB L2) per step on a concurrent stream, INSIDE the timed region
        (a serialised flush would have to drain the frame pipeline and time something the library never does)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        wall0 = time.perf_counter()
        ev0.record()
        chunk = batch if (batch > 1 and fn is step_device) else 1
        for k in range(0, steps, chunk):
            m = min(chunk, steps - k)
            if flush_l2:
                with torch.cuda.stream(flush_stream):
                    for _ in range(m):
                        flush.zero_()
            if chunk > 1:
                fn(m)
            else:
                fn()
        if finisher is not None:
            finisher()             # drain the pipeline inside the timed region
        ev1.record()
        barrier()
        wall = time.perf_counter() - wall0
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step_device()
    finish_device()
    launches0 = pt.KernelLaunches
    with ClockSampler(physical_gpu_index(local_rank)) as clocks:
        ms_total, _ = timed(step_device, args.steps, True, finish_device)
    launches = pt.KernelLaunches - launches0
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_total / args.steps, "note": "not a bench value"}), flush=True)
        pt.Dispose()
        return
    # back-to-back (image stays L2-resident between frames, as in the interactive app) and the end-to-end path
    ms_b2b, _ = timed(step_device, args.steps, False, finish_device)
    for _ in range(3):
        step_e2e()
    finish_e2e()
    e2e_steps = max(5, min(args.steps, 200))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    finish_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    # outside the timed region: the frame that reached host memory last is the image on the device, bit for bit
    e2e_verified = None
    try:
        if tiled is None:
            got = host_bufs[state["i"] & 1].numpy()
            want = pt.Result[..., :3]
            e2e_verified = bool((got.view(np.uint32) == want.view(np.uint32)).all())
        elif shared is not None:
            barrier()
            if rank == 0:
                got = shared.view(state["i"], (H, W, 3), torch.float32).numpy()
                want = tiled.flush()[..., :3].contiguous().cpu().numpy()
                e2e_verified = bool((got.view(np.uint32) == want.view(np.uint32)).all())
    except Exception as exc:      # noqa: BLE001
        e2e_verified = f"check failed to run: {exc}"

    # N = 1: the same end-to-end loop in the other two read-back formats (PCIe is the bound: 16 / 12 / 4 bytes per pixel)
    e2e_formats = None
    if tiled is None:
        e2e_formats = {}
        alt_buf = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
        for name, fmt in (("RGBA32F", ptb200.FORMAT_RGBA32F), ("RGBA8_display", ptb200.FORMAT_RGBA8)):
            def step_alt(fmt=fmt):
                pt.BasicDataUBO.SubData(64, 64, inv_view)
                pt.BasicDataUBO.SubData(128, 16, view_pos)
                pt.Render()
                pt.ReadResultAsync(alt_buf.data_ptr(), fmt)
            for _ in range(3):
                step_alt()
            pt.Synchronize()
            n_alt = max(5, min(args.steps, 100))
            t0 = time.perf_counter()
            for _ in range(n_alt):
                step_alt()
            pt.Synchronize()
            e2e_formats[name] = {"value": samples_per_step_() * n_alt / (time.perf_counter() - t0) / 1e6, "unit": "Msamples/s",
                                 "d2h_bytes_per_step": W * H * (16 if fmt == ptb200.FORMAT_RGBA32F else 4)}

    # kernel-only duration for the roofline (device time of the megakernel launches alone, max over ranks)
    pt.SetOverlap(1)                 # isolate the megakernel: one stream, no blend kernel, launches back to back
    pt.Render(3); pt.Synchronize()
    pt.Render(20)
    kern_ms = pt.LastRenderMs() / 20
    pt.SetOverlap(frames_in_flight)
    if world > 1:
        t = torch.tensor([kern_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kern_ms = float(t.item())

    samples_per_step = W * H * SPP
    value = samples_per_step * args.steps / (ms_total * 1e-3) / 1e6
    peak, peak_src = measured_peaks()
    algo_bytes = rows_local * W * 32            # 16 B load of the previous mean + 16 B store per pixel per dispatch
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    ncu = ncu_summary()
    line = {
        "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "flushed every step by a 160 MiB memset (> 126 MB L2) on a concurrent stream, inside the timed region",
                   "partition": f"interleaved {STRIPE_ROWS}-row stripes over {world} GPU(s); " + ("exchange fused into the blend kernel: peer stores into rank 0's image over NVLink (CUDA IPC), no NCCL on the data path" if fused else "one NCCL gather to rank 0 per frame + de-interleave, overlapped with the next frame's render") if world > 1 else "single GPU, no collective",
                   "kernel": f"persistent megakernel (ptb::megakernel) + blend kernel per frame, {frames_in_flight} frames in flight (ptb_set_overlap)"
                             + (f", {batch} frames per trace launch (PTB_BATCH)" if batch > 1 else "") + (f", grid / {grid_div} (PTB_GRID_DIV)" if grid_div > 1 else "")},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu.get("dram_bytes_per_launch"), "peak_source": peak_src, "kernel": "ptb::megakernel<false>",
                     "kernel_ms": kern_ms, "kernel_timing": "megakernel alone, in-place mode (ptb_set_overlap(1)), 20 launches back to back, CUDA events", "algorithmic_bytes_per_launch": algo_bytes,
                     "note": "the pass is FP32-issue-bound, not HBM-bound: ~3.4k lane-instructions per 32 B of image traffic (DESIGN.md); see issue_*",
                     "issue_active_pct": ncu.get("smsp_issue_active_pct"), "inst_executed_per_launch": ncu.get("inst_executed_per_launch"),
                     "issue": issue_roofline(ncu, kern_ms, world)},
        "e2e": {"value": samples_per_step * e2e_steps / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": (80 + 144) * world,
                "d2h_bytes_per_step": W * H * (12 if rgb_e2e else 16), "steps": e2e_steps, "format": "RGB32F" if rgb_e2e else "RGBA32F",
                "last_frame_on_host_equals_device_image": e2e_verified,
                "note": ("per step: InvView+ViewPos SubData (80 B host->library; the 144 B UBO rides in the kernel parameters), Render(), the frame read back to pinned host memory as RGB32F "
                         "(colour floats bit for bit; the constant alpha 1.0 of compute.glsl:129 is not shipped) through the pipelined read-back (snapshot/pack kernel on the render stream + copy stream, "
                         "overlapping the next Render()); one sync after the last step, inside the timed region. "
                         + ("" if world == 1 else ("N > 1: ptb_read_result_scatter_async — every rank writes its stripes into their rows of one full-frame image in pinned host memory shared by all ranks, each over its own PCIe link; the device-side exchange runs as in the device-timed loop"
                                                   if shared is not None else "N > 1 fallback: rank 0 copies the assembled RGBA32F image")))},
        "e2e_other_formats": e2e_formats,
        "back_to_back": {"value": samples_per_step * args.steps / (ms_b2b * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms_b2b / args.steps,
                         "note": "same steps without the concurrent L2 flush"},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
    }
    if rank == 0 and world == 1:
        line["cpu_baseline"] = cpu_baseline(pt)
        line["gl_proxy"] = gl_proxy()
    if tiled is not None and fused:
        tiled.exchange_ok()
    if shared is not None:
        pt.Synchronize()
        shared.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    pt.Dispose()


def cpu_baseline(pt):
    """The compiled reference shader (else the oracle port) on this box's host cores, a bounded sample of the same
    workload (rank 0, N = 1 only)."""
    import ptb200
    from oracle import oracle as O_
    O, kind, _ = cpu_reference()
    sc = ptb200.scene
    env = pt.ReadEnvironment()
    scene, cam = sc.load_default_scene(), sc.default_camera()
    basic, ubo = sc.basic_data_bytes(cam, W, H), scene.ubo_bytes()
    img = np.zeros((H, W, 4), np.float32)
    threads = host_threads()
    kw = dict(spp=SPP, ray_depth=RAY_DEPTH, focal_length=FOCAL, aperture_diameter=APERTURE, n_spheres=48, n_cuboids=7, n_threads=threads)
    O.render(img, basic, ubo, env, frame=0, **kw)                 # warm-up / thread pool start
    frames, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 10.0 and frames < 64:
        frames += 1
        O.render(img, basic, ubo, env, frame=frames, **kw)
    dt = time.perf_counter() - t0
    return {"value": W * H * SPP * frames / dt / 1e6, "unit": "Msamples/s", "cores": threads, "kind": kind,
            "sample": f"{frames} full 1920x1080 frames at SPP 1 ({dt:.1f} s), OpenMP over rows"}


def gl_proxy():
    """The reference's compute.glsl compiled by nvcc and dispatched in the reference's launch shape on this GPU
    (tools/gl_proxy_probe.py; oracle/build_ref.py --cuda).  Runs in a subprocess after all timed regions: whatever happens
    there cannot disturb the bench.  Reported next to our numbers; not a parity path, not the optimisation target."""
    import subprocess
    from oracle import build_ref
    if not (os.path.exists(build_ref.cuda_lib(False)) or os.path.exists(build_ref.cuda_lib(True))):
        return {"unavailable": "oracle/_ref/libglsl_ref_cuda*.so not built (python oracle/build_ref.py --cuda [--fast], needs /root/reference)"}
    try:
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gl_proxy_probe.py"), "--frames", "20"], capture_output=True,
                             text=True, timeout=120)
        lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        if res.returncode != 0 or not lines:
            return {"error": f"probe exited with {res.returncode}: {res.stderr.strip()[-300:]}"}
        return json.loads(lines[-1])
    except Exception as exc:      # noqa: BLE001
        return {"error": str(exc)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--profile", action="store_true", help="profiling aid: only warm-up + timed steps (for runs under ncu); prints no bench line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        import torch
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    elif args.gpus > 1:
        print(json.dumps({"error": "launch with torch.distributed.run --nproc-per-node N for --gpus N > 1"}))
        sys.exit(2)
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
