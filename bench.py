#!/usr/bin/env python
"""bench.py — Msamples/s of the path-tracing pass (BASELINE.json's metric) on one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4] [--precision auto|fast|exact]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workloads (BASELINE.json `configs`):  c2 (default) = default demo scene, 1920x1080, progressive accumulation at SPP 1 per
frame — the configuration the metric is quoted on;  c3 = synthetic 1024 spheres + 256 cuboids, 1080p, rayDepth 8;
c4 = the default scene at 3840x2160 (the 8-GPU tile-partition case).

A step = one PathTracer.Render() of the whole frame (one dispatch in the reference, PathTracer.cs:114-129) = W*H*SPP samples.
Consecutive steps are consecutive frames of ONE progressive render (static camera), which is what configs[1] describes; the
library traces them in batches of up to 16 frames per persistent-kernel launch (ptb_set_batch) and folds each batch into the
accumulation image with one blend kernel — bit-identical to frame-by-frame dispatches.  At N > 1 the frame is cut into
interleaved 8-row stripes, one process per GPU, and every frame still reaches rank 0 through the path's one exchange (fused
into the blend kernel: peer stores over NVLink; PTB_EXCHANGE=nccl selects an NCCL gather).  Fixed total work => "strong".

Arithmetic: `value` is measured with the precision named in config.precision.  `--precision auto` (default) first checks
the fast build (MUFU + FMA) against the exact build on this very workload at matched seeds — per-channel MSE of the
accumulated image must be below the north star's 1e-6 — and uses it when it passes; the exact build (bit-identical to the
CPU oracle and the compiled reference shaders) is timed beside it and reported under `exact`.

Prints ONE JSON line (rank 0).  `value`: CUDA events on the launching stream around K steps, inputs resident in HBM.
L2: no flush kernel runs inside the timed region — the frame estimates rotate through 3 x 16 scratch images per rank
(1.6 GB at 1080p on one GPU, 200 MB on an eighth of the frame), far more than the 126 MB L2, so no step finds its inputs or outputs cached.
`e2e`: the same metric through the public API with host buffers, one Render() per step: the per-frame camera UBO upload
the C# host does (MainWindow.cs:131-132) and a read-back of every frame into pinned host memory as RGB32F.
`--impl reference` times the reference's own compute shader on the host CPUs (oracle/_ref: compute.glsl compiled by g++ from
/root/reference, kind "reference"; else the oracle port), all host threads, a bounded sample per step.
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

SPP, FOCAL, APERTURE = 1, 20.0, 0.14
CONFIGS = {
    "c2": dict(W=1920, H=1080, ray_depth=13, scene="default", gate_frames=1024,
               workload="default demo scene (48 spheres + 7 cuboids), 1920x1080, SPP 1 per frame, rayDepth 13, 256^2 atmosphere env (BASELINE configs[1])"),
    "c3": dict(W=1920, H=1080, ray_depth=8, scene="synthetic", gate_frames=128,
               workload="synthetic scene (1024 spheres + 256 cuboids, seed 1234), 1920x1080, SPP 1 per frame, rayDepth 8, 256^2 atmosphere env (BASELINE configs[2])"),
    "c4": dict(W=3840, H=2160, ray_depth=13, scene="default", gate_frames=256,
               workload="default demo scene (48 spheres + 7 cuboids), 3840x2160, SPP 1 per frame, rayDepth 13, 256^2 atmosphere env (BASELINE configs[3])"),
}
STRIPE_ROWS = 8
BATCH = max(1, min(16, int(os.environ.get("PTB_BATCH", "16"))))

ROTATE_ROOTS = os.environ.get("PTB_ROOTS", "rotate") != "rank0"   # frame q is assembled on rank q % N (one gather per frame; no single GPU's NVLink ingress carries them all)
EXCHANGE_SLOTS = int(os.environ.get("PTB_SLOTS", "8" if ROTATE_ROOTS else "32"))   # full-image buffers per root (>= 2 batches of frames in flight over all roots)
MSE_TOLERANCE = 1e-6                                         # north star: per-channel MSE at matched seed


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_summary():
    """Figures of the committed `ncu --set full` capture of the megakernel (profiles/ncu_summary.json), if any."""
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
            self._stop.wait(0.005)

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


def load_scene(cfg):
    import ptb200
    sc = ptb200.scene
    return sc.load_default_scene() if cfg["scene"] == "default" else sc.synthetic_scene(1024, 256)


# ================================================================================================ reference arm
def host_threads():
    """Threads the CPU arm may use: the cores this process is allowed on.  torchrun exports OMP_NUM_THREADS=1 to its workers,
    which would silently turn the reference arm into a single-core run at N > 1 — the count is passed explicitly instead."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference(max_spheres=256):
    """(module, kind, description): the compiled reference shaders when oracle/_ref was built, else the oracle port."""
    from oracle import oracle as O, ref as R
    if max_spheres == 256 and R.available():
        return R, "reference", "the reference's compute.glsl compiled for the CPU (g++, oracle/build_ref.py -> oracle/_ref/libglsl_ref.so); its C# + OpenGL host needs .NET + GL 4.5, absent here"
    big = os.path.join(ROOT, "oracle", "_ref", f"libglsl_ref_{max_spheres}x256.so")
    if max_spheres != 256 and os.path.exists(big):
        return R.variant(big), "reference", f"the reference's compute.glsl with its two UBO array lengths rewritten to {max_spheres}/256 (build_ref.py --capacity), compiled for the CPU"
    return O, "port", "CPU oracle (C restatement of compute.glsl); oracle/_ref not built on this machine"


def run_reference(args, rank, world):
    """The reference's shader on the host CPUs (oracle/_ref, kind "reference"; else the oracle, kind "port"), all threads,
    a bounded sample per step."""
    if rank != 0:
        return
    import ptb200
    from oracle import oracle as O_
    cfg = CONFIGS[args.config]
    W, H = cfg["W"], cfg["H"]
    scene = load_scene(cfg)
    O, kind, kind_text = cpu_reference(scene.max_spheres)
    sc = ptb200.scene
    threads = host_threads()
    env = O_.atmosphere(256, sc.atmosphere_ubo_bytes(), sc.atmosphere_light_pos(0.5), 15.0, 50, 15, threads)
    cam = sc.default_camera()
    basic, ubo = sc.basic_data_bytes(cam, W, H), scene.ubo_bytes()
    img = np.zeros((H, W, 4), np.float32)
    kw = dict(spp=SPP, ray_depth=cfg["ray_depth"], focal_length=FOCAL, aperture_diameter=APERTURE, n_spheres=len(scene.spheres),
              n_cuboids=len(scene.cuboids), max_spheres=scene.max_spheres, n_threads=threads)
    # calibrate on a sparse sample, then bound the per-step sample so warmup + steps stay under ~150 s
    t = time.perf_counter()
    O.render(img, basic, ubo, env, frame=0, y_step=16, **kw)
    full_s = (time.perf_counter() - t) * 16
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
    sample = f"every {y_step}-th row of the {W}x{H} frame per step ({rows_per_step} rows, {px_per_step} samples/step)" if y_step > 1 else f"the full {W}x{H} frame per step"
    line = {"impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "name": args.config, "reference_kind": kind_text},
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

    cfg = CONFIGS[args.config]
    W, H, RAY_DEPTH = cfg["W"], cfg["H"], cfg["ray_depth"]
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sc = ptb200.scene
    scene, cam = load_scene(cfg), sc.default_camera()
    pt = ptb200.PathTracer(None, W, H, RAY_DEPTH, SPP, FOCAL, APERTURE, max_spheres=scene.max_spheres, max_cuboids=scene.max_cuboids, device=local_rank)
    pt.SetStream(torch.cuda.current_stream(dev).cuda_stream)
    frames_in_flight = int(os.environ.get("PTB_OVERLAP", "2"))
    pt.SetOverlap(frames_in_flight)
    pt.SetBatch(BATCH)
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
                tiled = D.TiledPathTracer(pt, rank, world, STRIPE_ROWS, device=dev, fused=True, slots=EXCHANGE_SLOTS, rgb=True, rotate=ROTATE_ROOTS)
                ok = torch.ones(1, device=dev)
            except Exception as exc:      # noqa: BLE001
                print(f"rank {rank}: fused exchange unavailable ({exc}); falling back to the NCCL gather", file=sys.stderr, flush=True)
                ok = torch.zeros(1, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() < 1:
                fused, tiled = False, None
        if tiled is None:
            tiled = D.TiledPathTracer(pt, rank, world, STRIPE_ROWS, device=dev, fused=False)
    inv_view = sc.matrix_bytes(sc.inverted(cam.View))
    view_pos = np.append(np.asarray(cam.Position, np.float32), np.float32(0)).tobytes()
    rows_local = pt.Result.shape[0]
    samples_per_step = W * H * SPP

    t_start = time.perf_counter()

    def log(msg):
        if os.environ.get("PTB_BENCH_LOG"):
            print(f"[bench rank {rank} +{time.perf_counter() - t_start:6.1f}s] {msg}", file=sys.stderr, flush=True)

    def local_image():
        ptr, _ = pt.ResultDevicePtr()
        return torch.as_tensor(D._DeviceBuffer(ptr, (rows_local, W, 4)), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device(n=1):
        if tiled is None:
            pt.Render(n)
        elif fused:
            tiled.step_batch(n)    # one trace launch per batch; every frame still reaches rank 0 through its own slot
        else:
            for _ in range(n):
                tiled.step()       # render f; finish gather f-1 (it overlapped this render); start gather f

    def finish_device():
        if tiled is not None:
            tiled.flush()

    def run_steps(steps):
        for k in range(0, steps, BATCH):
            step_device(min(BATCH, steps - k))
        finish_device()

    # ------------------------------------------------------------------ precision gate (outside every timed region)
    gate = None
    precision = args.precision
    if precision in ("auto", "fast") and not (args.no_gate and precision == "fast"):
        n_gate = cfg["gate_frames"]
        imgs = {}
        for prec in (ptb200.PRECISION_EXACT, ptb200.PRECISION_FAST):
            log(f"gate: precision {prec}, {n_gate} frames")
            pt.SetPrecision(prec)
            pt.ResetRenderer()
            run_steps(n_gate)
            pt.Synchronize()
            torch.cuda.synchronize(dev)
            imgs[prec] = local_image()[..., :3].double().clone()
        diff = imgs[ptb200.PRECISION_EXACT] - imgs[ptb200.PRECISION_FAST]
        stats = torch.stack([(diff ** 2).sum(dim=(0, 1)), torch.full((3,), float(diff[..., 0].numel()), device=dev, dtype=torch.float64)])
        finite = torch.isfinite(diff).all().float()
        if world > 1:
            dist.all_reduce(stats)
            dist.all_reduce(finite, op=dist.ReduceOp.MIN)
        mse = (stats[0] / stats[1]).tolist()
        passed = bool(finite.item() >= 1 and max(mse) < MSE_TOLERANCE)
        gate = {"frames": n_gate, "per_channel_mse_fast_vs_exact": mse, "tolerance": MSE_TOLERANCE, "passed": passed,
                "note": "accumulated image after `frames` frames at SPP 1, matched seeds, this workload, all ranks' stripes"}
        del imgs, diff
        if precision == "auto":
            precision = "fast" if passed else "exact"
    headline = ptb200.PRECISION_FAST if precision == "fast" else ptb200.PRECISION_EXACT

    # ------------------------------------------------------------------ device-timed region
    def timed(steps):
        """K steps enqueued in batches on the launching stream, one CUDA-event bracket around the whole region (the pipeline is
        drained inside it).  No flush kernel: the estimates rotate through 3 x BATCH scratch images (> L2)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        run_steps(steps)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def measure(prec, steps, with_clocks):
        pt.SetPrecision(prec)
        pt.ResetRenderer()
        run_steps(max(args.warmup, 3))
        launches0 = pt.KernelLaunches
        if with_clocks:
            with ClockSampler(physical_gpu_index(local_rank)) as clocks:
                ms = timed(steps)
            clk = clocks.summary()
        else:
            ms, clk = timed(steps), None
        launches = pt.KernelLaunches - launches0
        # the megakernel launches themselves, in the same batched / pipelined / tiled mode (device-side bracket of each launch)
        barrier()
        pt.SetKernelTiming(True)
        run_steps(max(BATCH, min(steps, 8 * BATCH)))
        barrier()
        kt = pt.KernelTime()
        pt.SetKernelTiming(False)
        kern = torch.tensor([kt["ms"] / max(1, kt["frames"]), kt["ms"] / max(1, kt["launches"])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(kern, op=dist.ReduceOp.MAX)
        return ms, launches, clk, float(kern[0].item()), float(kern[1].item()), kt["frames"] // max(1, kt["launches"])

    log("timed region (headline precision)")
    ms_total, launches, clocks, kern_ms, kern_launch_ms, frames_per_launch = measure(headline, args.steps, True)
    log(f"timed region done: {ms_total / args.steps:.4f} ms/step")
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_total / args.steps, "note": "not a bench value"}), flush=True)
        pt.Dispose()
        return
    exact = None
    if headline != ptb200.PRECISION_EXACT:
        n_x = max(BATCH, args.steps // 2)
        ms_x, _, _, kern_x, _, _ = measure(ptb200.PRECISION_EXACT, n_x, False)
        exact = {"value": samples_per_step * n_x / (ms_x * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms_x / n_x, "kernel_ms": kern_x,
                 "note": "the exact build (bit-identical to the CPU oracle and the compiled reference shaders), same mode, half the steps"}
        pt.SetPrecision(headline)

    # ------------------------------------------------------------------ end to end
    shared = None
    if world > 1 and os.environ.get("PTB_E2E", "scatter") == "scatter":
        try:
            # NUMA placement: every rank first-touches the pages of its own stripes from the CPUs next to its GPU
            shared = D.SharedHostFrame(W * H * 12, 2, rank, world, stripe_bytes=STRIPE_ROWS * W * 12,
                                       numa_cpus=D.gpu_numa_cpus(physical_gpu_index(local_rank)))
        except Exception as exc:      # noqa: BLE001  (collective: raised on every rank or on none)
            if rank == 0:
                print(f"shared host frame unavailable ({exc}); e2e falls back to rank 0's copy of the assembled image", file=sys.stderr, flush=True)
            shared = None
    rgb_e2e = world == 1 or shared is not None
    every_rank_roots = tiled is not None and fused and tiled.rotate
    host_bufs = [torch.empty((H if (rank == 0 or (every_rank_roots and shared is None)) else 1, W, 3 if tiled is None else tiled.channels), dtype=torch.float32).pin_memory() for _ in range(2)]
    side = torch.cuda.Stream(device=dev)
    copy_done = [torch.cuda.Event(), torch.cuda.Event()]
    snap = [torch.empty((H, W, tiled.channels), dtype=torch.float32, device=dev) for _ in range(2)] if ((rank == 0 or every_rank_roots) and world > 1 and shared is None) else None
    state = {"i": 0}

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
        elif fused:
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
            tiled.step_fused(consumer=consume if snap is not None else None)
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

    log("e2e warm-up")
    for _ in range(3):
        step_e2e()
    finish_e2e()
    log("e2e timed")
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
            root = tiled.last_root() if fused else 0          # the rank the last frame was assembled on (rotating roots)
            dev_img = torch.empty((H, W, tiled.channels), dtype=torch.float32, device=dev)
            if rank == root:
                dev_img.copy_(tiled.flush())
            dist.broadcast(dev_img, src=root)
            if rank == 0:
                got = shared.view(state["i"], (H, W, 3), torch.float32).numpy()
                want = dev_img[..., :3].contiguous().cpu().numpy()
                e2e_verified = bool((got.view(np.uint32) == want.view(np.uint32)).all())
            del dev_img
    except Exception as exc:      # noqa: BLE001
        e2e_verified = f"check failed to run: {exc}"

    # what this box's PCIe link gives a plain pinned device->host copy (outside every timed region): the ceiling of
    # the e2e number, and it varies between boxes of the pool (57 GB/s on most, ~20 GB/s seen on one)
    pcie_gbps = None
    if rank == 0:
        try:
            probe_d = torch.empty(128 << 20, dtype=torch.uint8, device=dev)
            probe_h = torch.empty(128 << 20, dtype=torch.uint8).pin_memory()
            probe_h.copy_(probe_d, non_blocking=True); torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(4):
                probe_h.copy_(probe_d, non_blocking=True)
            torch.cuda.synchronize(dev)
            pcie_gbps = 4 * (128 << 20) / (time.perf_counter() - t0) / 1e9
            del probe_d, probe_h
        except Exception:      # noqa: BLE001
            pcie_gbps = None

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
            e2e_formats[name] = {"value": samples_per_step * n_alt / (time.perf_counter() - t0) / 1e6, "unit": "Msamples/s",
                                 "d2h_bytes_per_step": W * H * (16 if fmt == ptb200.FORMAT_RGBA32F else 4)}

        # A host that does not look at every frame: one Render(BATCH) call and ONE lossless read-back per call.  At
        # thousands of frames per second no display consumes each frame (the reference's window shows at most one per
        # monitor refresh, MainWindow.cs:44-56 under GameWindow.Run); informational, the `e2e` key stays one read per frame.
        def step_batch_read():
            pt.BasicDataUBO.SubData(64, 64, inv_view)
            pt.BasicDataUBO.SubData(128, 16, view_pos)
            pt.Render(BATCH)
            pt.ReadResultAsync(host_bufs[0].data_ptr(), ptb200.FORMAT_RGB32F)
        for _ in range(3):
            step_batch_read()
        pt.Synchronize()
        n_alt = max(4, min(args.steps // BATCH, 40))
        t0 = time.perf_counter()
        for _ in range(n_alt):
            step_batch_read()
        pt.Synchronize()
        e2e_formats[f"RGB32F_one_readback_per_{BATCH}_frames"] = {
            "value": samples_per_step * BATCH * n_alt / (time.perf_counter() - t0) / 1e6, "unit": "Msamples/s", "d2h_bytes_per_step": W * H * 12 / BATCH,
            "note": f"one Render({BATCH}) call + one RGB32F read-back per call: a host that displays every {BATCH}th frame"}

    # ------------------------------------------------------------------ N > 1: the exchanged frame against ONE GPU rendering alone
    exchange_check = None
    log("e2e done; exchange check")
    if tiled is not None:
        n_chk = 8
        pt.Synchronize()
        pt.ResetRenderer()
        barrier()
        keep = {}

        def grab(full):
            keep["img"] = full            # the slot stays valid: nothing is rendered after these frames
        if fused:
            tiled.step_batch(n_chk, consumer=grab)
            root = tiled.last_root()      # the rank the last frame was assembled on
        else:
            for _ in range(n_chk):
                tiled.step()
            keep["img"] = tiled.flush()
            root = 0
        barrier()
        got = torch.empty((H, W, tiled.channels), dtype=torch.float32, device=dev)
        if rank == root:
            got.copy_(keep["img"])
        dist.broadcast(got, src=root)     # to rank 0 for the comparison (outside every timed region)
        if rank == 0:
            solo = ptb200.PathTracer(None, W, H, RAY_DEPTH, SPP, FOCAL, APERTURE, max_spheres=scene.max_spheres, max_cuboids=scene.max_cuboids, device=local_rank)
            solo.SetPrecision(headline)
            solo.GenerateAtmosphere(256, 50, 15, 0.5, 15.0); solo.LoadScene(scene); solo.SetCamera(cam)
            solo.Render(n_chk); solo.Synchronize()
            ptr, _ = solo.ResultDevicePtr()
            want = torch.as_tensor(D._DeviceBuffer(ptr, (H, W, 4)), device=dev)
            want = want[..., :got.shape[-1]]         # RGB32F slots carry the three colour floats; alpha is the constant 1.0
            same = bool((got.view(torch.int32) == want.contiguous().view(torch.int32)).all().item())
            exchange_check = {"exchange_equals_single_gpu": same, "frames": n_chk, "assembled_on_rank": root,
                              "max_abs_diff": float((got - want).abs().max().item()),
                              "note": f"{n_chk} frames from a reset through the {world}-GPU exchange vs the same frames rendered by rank 0's GPU alone (untiled), compared on the device"}
            del want
            solo.Dispose()
        barrier()

    log("exchange check done")
    value = samples_per_step * args.steps / (ms_total * 1e-3) / 1e6
    peak, peak_src = measured_peaks()
    algo_bytes = rows_local * W * 32 * frames_per_launch   # per pixel and frame: 16 B load of the previous mean + 16 B store
    achieved = algo_bytes / (kern_launch_ms * 1e-3) / 1e9
    ncu = ncu_summary() if (world == 1 and args.config == "c2") else {}
    prec_name = "fast" if headline == ptb200.PRECISION_FAST else "exact"
    # The bound that actually limits the pass (DESIGN.md §4 Roofline): warp-instructions issued per second against
    # SMs x 4 schedulers x 1 instruction/clock at the clock sampled during the timed region.  The instruction count is the
    # committed ncu capture's (same kernel, same launch shape: deterministic work), the duration is this run's live figure.
    issue = None
    if ncu.get("inst_executed_per_launch") and prec_name == "fast" and ncu.get("frames_per_launch") == frames_per_launch and kern_launch_ms:
        sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
        issue_peak = torch.cuda.get_device_properties(dev).multi_processor_count * 4 * sm_mhz * 1e6
        issue_rate = float(ncu["inst_executed_per_launch"]) / (kern_launch_ms * 1e-3)
        issue = {"achieved": issue_rate / 1e12, "peak": issue_peak / 1e12, "unit": "T warp-instructions/s", "frac": issue_rate / issue_peak,
                 "note": "instructions per launch from the committed ncu capture / this run's live launch duration; peak = SMs x 4 schedulers x SM clock under load"}
    line = {
        "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": cfg["workload"], "name": args.config,
                   "precision": (f"{prec_name}: " + ("MUFU rcp/rsq/sin/cos/ex2 + FMA contraction (ptb_set_precision), within the north star's per-channel MSE < 1e-6 of the exact build on this workload (see precision_gate)"
                                                     if prec_name == "fast" else "the evaluation model of DESIGN.md §2, bit-identical to the CPU oracle and the compiled reference shaders")),
                   "l2": f"no flush kernel in the timed region: the frame estimates rotate through 3 x {BATCH} scratch images per rank ({3 * BATCH * rows_local * W * 16 / 1e6:.0f} MB on this rank, L2 is 126 MB); inputs larger than L2",
                   "partition": (f"interleaved {STRIPE_ROWS}-row stripes over {world} GPU(s); " + (f"exchange fused into the batch blend kernel: peer stores into the frame's root over NVLink (CUDA IPC; " + ("rotating roots: frame q is assembled on rank q % N" if ROTATE_ROOTS else "every frame on rank 0") + f"; {EXCHANGE_SLOTS} RGB32F slots per root: colour floats bit for bit, the constant alpha is not shipped), no NCCL on the data path" if fused else "one NCCL gather to rank 0 per frame + de-interleave, overlapped with the next frame's render")) if world > 1 else "single GPU, no collective",
                   "kernel": f"persistent megakernel, {BATCH} consecutive frames per launch (ptb_set_batch) + one blend kernel per batch, {frames_in_flight} batches in flight; ray-classification table for scenes of <= 64 primitives, shared-memory BVH above 96"},
        "precision_gate": gate,
        "exact": exact,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu.get("dram_bytes_per_launch"), "peak_source": peak_src, "kernel": f"ptb{'_fast' if prec_name == 'fast' else ''}::megakernel",
                     "kernel_ms": kern_ms, "kernel_ms_per_launch": kern_launch_ms, "frames_per_launch": frames_per_launch,
                     "kernel_timing": "device-side bracket of every megakernel launch (first CTA start -> last CTA end, %globaltimer) in the same batched / pipelined / tiled mode as `value` (ptb_set_kernel_timing), max over ranks; kernel_ms is per frame; consecutive launches overlap in the drain of the previous grid, so kernel_ms can exceed ms_per_step by that overlap",
                     "algorithmic_bytes_per_launch": algo_bytes,
                     "note": "the pass is FP32-issue-bound, not HBM-bound: thousands of lane-instructions per 32 B of image traffic (DESIGN.md); `traffic` and ncu_capture come from the committed single-GPU c2 capture named in profiles/ncu_summary.json and are attached to that configuration only",
                     "ncu_capture": ({k: ncu.get(k) for k in ("source", "kernel", "smsp_issue_active_pct", "inst_executed_per_launch", "thread_inst_per_sample", "frames_per_launch")} if ncu else None),
                     "issue": issue},
        "e2e": {"value": samples_per_step * e2e_steps / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": (80 + 144) * world,
                "d2h_bytes_per_step": W * H * (12 if (rgb_e2e or (tiled is not None and tiled.channels == 3)) else 16), "steps": e2e_steps,
                "format": "RGB32F" if (rgb_e2e or (tiled is not None and tiled.channels == 3)) else "RGBA32F",
                "last_frame_on_host_equals_device_image": e2e_verified,
                "pcie_d2h_gbps": pcie_gbps,
                "host_frame_numa": getattr(shared, "numa", None) if shared is not None else None,
                "note": ("one Render() per step (no batching: every frame is read back): InvView+ViewPos SubData (80 B host->library; the 144 B UBO rides in the kernel parameters), Render(), the frame read back to pinned host memory as RGB32F "
                         "(colour floats bit for bit; the constant alpha 1.0 of compute.glsl:129 is not shipped) through the pipelined read-back (snapshot/pack kernel on the render stream + copy stream, "
                         "overlapping the next Render()); one sync after the last step, inside the timed region; PCIe-bound. "
                         + ("" if world == 1 else ("N > 1: ptb_read_result_scatter_async — every rank writes its stripes into their rows of one full-frame image in pinned host memory shared by all ranks, each over its own PCIe link; the device-side exchange runs as in the device-timed loop"
                                                   if shared is not None else "N > 1 fallback: rank 0 copies the assembled RGBA32F image")))},
        "e2e_other_formats": e2e_formats,
        "exchange_check": exchange_check,
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if rank == 0 and world == 1:
        line["cpu_baseline"] = cpu_baseline(pt, cfg, scene)
        if args.config == "c2":
            line["gl_proxy"] = gl_proxy()
    if tiled is not None and fused:
        tiled.exchange_ok()
    if shared is not None:
        pt.Synchronize()
        shared.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    pt.Dispose()


def cpu_baseline(pt, cfg, scene):
    """The compiled reference shader (else the oracle port) on this box's host cores, a bounded sample of the same
    workload (rank 0, N = 1 only)."""
    import ptb200
    O, kind, _ = cpu_reference(scene.max_spheres)
    sc = ptb200.scene
    W, H = cfg["W"], cfg["H"]
    env = pt.ReadEnvironment()
    cam = sc.default_camera()
    basic, ubo = sc.basic_data_bytes(cam, W, H), scene.ubo_bytes()
    img = np.zeros((H, W, 4), np.float32)
    threads = host_threads()
    kw = dict(spp=SPP, ray_depth=cfg["ray_depth"], focal_length=FOCAL, aperture_diameter=APERTURE, n_spheres=len(scene.spheres),
              n_cuboids=len(scene.cuboids), max_spheres=scene.max_spheres, n_threads=threads)
    y_step = 1 if cfg["scene"] == "default" and W <= 1920 else 8      # the large configs take a bounded sample: every 8th row
    O.render(img, basic, ubo, env, frame=0, y_step=max(y_step, 4), **kw)                 # warm-up / thread pool start
    frames, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 10.0 and frames < 64:
        frames += 1
        O.render(img, basic, ubo, env, frame=frames, y_step=y_step, **kw)
    dt = time.perf_counter() - t0
    rows = len(range(0, H, y_step))
    return {"value": W * rows * SPP * frames / dt / 1e6, "unit": "Msamples/s", "cores": threads, "kind": kind,
            "sample": f"{frames} passes over " + ("the full frame" if y_step == 1 else f"every {y_step}-th row of the frame") + f" ({W}x{rows} pixels each) at SPP 1 ({dt:.1f} s), OpenMP over rows"}


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
    ap.add_argument("--steps", type=int, default=320)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default="auto", choices=["auto", "fast", "exact"])
    ap.add_argument("--no-gate", action="store_true", help="profiling aid: skip the fast-vs-exact check (with --precision fast|exact); a bench line without the gate says so")
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
