"""Pixel-tile partition across GPUs and the per-frame gather (one process per GPU, torch.distributed plumbing).

The reference has no multi-GPU code.  Pixels are independent — each reads only its own previous
value and a seed that is a function of the *global* pixel coordinate (compute.glsl:104-106,126-129) — so the
image is cut into interleaved stripes of `stripe_rows` rows, stripe s belonging to rank s % world.  Interleaving
(rather than one contiguous band per rank) balances sky rows (1 bounce) against interior rows (many bounces).
Every rank keeps its stripes resident (the progressive mean is local); one collective per frame — a gather of the
rank-major stripe buffers to rank 0 — followed by a de-interleave rebuilds the row-major image.

The partition arithmetic here is the host mirror of compute_local_rows()/global_row() in csrc; tests check the two
against each other and, on CPU with gloo at world_size 2, the whole gather path against an unpartitioned render.
"""
from __future__ import annotations

import numpy as np

DEFAULT_STRIPE_ROWS = 8


def local_rows_of(rank: int, world: int, stripe_rows: int, height: int) -> np.ndarray:
    """Global row index of every local row of `rank`, in local order."""
    rows = []
    n_stripes = (height + stripe_rows - 1) // stripe_rows
    for s in range(rank, n_stripes, world):
        rows.extend(range(s * stripe_rows, min((s + 1) * stripe_rows, height)))
    return np.asarray(rows, dtype=np.int64)


def local_row_count(rank: int, world: int, stripe_rows: int, height: int) -> int:
    return int(local_rows_of(rank, world, stripe_rows, height).size)


def max_local_rows(world: int, stripe_rows: int, height: int) -> int:
    """Rank 0 always holds the most rows; every rank's buffer is padded to this so the gather is uniform."""
    return local_row_count(0, world, stripe_rows, height)


def deinterleave_host(gathered: np.ndarray, height: int, world: int, stripe_rows: int) -> np.ndarray:
    """gathered: (world, max_local_rows, W, C) rank-major stripe buffers -> (height, W, C) row-major image."""
    out = np.empty((height,) + gathered.shape[2:], dtype=gathered.dtype)
    for r in range(world):
        rows = local_rows_of(r, world, stripe_rows, height)
        out[rows] = gathered[r, :rows.size]
    return out


def scatter_plan(rank: int, world: int, stripe_rows: int, height: int):
    """Host mirror of ptb_read_result_scatter_async's copy plan, in ROWS: (first local row, first global row, rows per chunk,
    chunks, global row pitch) for the strided copy of the full stripes, and (first local row, first global row, rows) for the
    ragged last stripe (rows == 0 if this rank does not own one)."""
    n_local = local_row_count(rank, world, stripe_rows, height)
    n_full, tail = divmod(n_local, stripe_rows)
    strided = (0, rank * stripe_rows, stripe_rows, n_full, world * stripe_rows)
    ragged = (n_full * stripe_rows, (rank + n_full * world) * stripe_rows, tail)
    return strided, ragged


def scatter_rows_host(local: np.ndarray, frame: np.ndarray, rank: int, world: int, stripe_rows: int) -> None:
    """What the scatter read-back does, on host arrays: local (compact stripes, >= local rows) -> rows of `frame`."""
    (l0, g0, rows, chunks, pitch), (tl, tg, tail) = scatter_plan(rank, world, stripe_rows, frame.shape[0])
    for k in range(chunks):
        frame[g0 + k * pitch:g0 + k * pitch + rows] = local[l0 + k * rows:l0 + (k + 1) * rows]
    if tail:
        frame[tg:tg + tail] = local[tl:tl + tail]


def gather_stripes(local, world: int, dst: int = 0, group=None, async_op: bool = False):
    """One collective: gather every rank's (max_local_rows, W, 4) stripe buffer to `dst`.
    `local` is a torch tensor (CUDA with nccl, CPU with gloo).  Returns (gathered or None, work or None)."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    if world == 1:
        return local.unsqueeze(0), None
    if rank == dst:
        gathered = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
        work = dist.gather(local, list(gathered.unbind(0)), dst=dst, group=group, async_op=async_op)
        return gathered, work
    work = dist.gather(local, None, dst=dst, group=group, async_op=async_op)
    return None, work


class _DeviceBuffer:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr: int, shape, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class TiledPathTracer:
    """One rank's share of a frame: a PathTracer restricted to this rank's stripes + the per-frame gather.

    Usage (one process per GPU, torch.distributed initialised with nccl):
        tp = TiledPathTracer(tracer, rank, world)
        tp.render(); full = tp.gather()              # simple: full image on rank 0, in stream order
        for f in range(n): tp.step()                  # pipelined: the gather of frame f overlaps the render of frame f+1
        full = tp.flush()
    """

    def __init__(self, tracer, rank: int, world: int, stripe_rows: int = DEFAULT_STRIPE_ROWS, device=None, fused: bool = False,
                 slots: int = 2, rgb: bool = False, rotate: bool = False):
        import torch

        self.tracer, self.rank, self.world, self.stripe_rows = tracer, rank, world, stripe_rows
        self.fused = bool(fused) and world > 1
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        # kernels and collectives share torch's current stream, so stream order is the only synchronisation needed
        tracer.SetStream(torch.cuda.current_stream(self.device).cuda_stream)
        tracer.SetTile(rank, world, stripe_rows)
        self.width, self.height = tracer.Width, tracer.Height
        self.max_rows = max_local_rows(world, stripe_rows, self.height)
        ptr, nbytes = tracer.ResultDevicePtr()
        assert nbytes >= self.max_rows * self.width * 16
        self._holder = _DeviceBuffer(ptr, (self.max_rows, self.width, 4))
        self.local = torch.as_tensor(self._holder, device=self.device)
        # two of everything that a gather in flight reads or writes, so frame f+1 never waits for frame f's exchange
        self.staging = [torch.empty_like(self.local) for _ in range(2)]
        self.gathered = [torch.empty((world,) + tuple(self.local.shape), dtype=torch.float32, device=self.device) if rank == 0 else None
                         for _ in range(2)]
        self.full = [torch.empty((self.height, self.width, 4), dtype=torch.float32, device=self.device) if rank == 0 else None
                     for _ in range(2)]
        self._k = 0
        self._pending = None       # (buffer index, work)
        self._last_full = None
        self._slot_tensors = {}
        self.channels = 3 if (rgb and self.fused) else 4       # fused exchange: RGB32F slots ship 12 instead of 16 bytes per pixel
        self.rotate = bool(rotate) and self.fused             # frame q is assembled on rank q % world instead of always on rank 0
        if self.fused:
            self._init_fused(slots)

    # ------------------------------------------------------------------ fused exchange (peer stores instead of NCCL)
    def _init_fused(self, slots: int) -> None:
        """ptb_exchange_*: a root allocates [flags | slots x full image]; its CUDA-IPC handle (64 bytes) travels through
        torch.distributed and is mapped by the other ranks, whose blend kernels then store straight into the root's image.
        One root (rank 0) by default; with `rotate` every rank is the root of every world-th frame and maps all the others."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib

        L, ctx = self.tracer._L, self.tracer._ctx
        self.slots = int(slots)
        fmt = 1 if self.channels == 3 else 0
        if self.rotate:
            _lib.check(L.ptb_exchange_init_roots(ctx, slots, fmt, 1))
            raw = C.create_string_buffer(64)
            _lib.check(L.ptb_exchange_handle(ctx, raw))
            mine = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).to(self.device)
            handles = [torch.zeros(64, dtype=torch.uint8, device=self.device) for _ in range(self.world)]
            dist.all_gather(handles, mine)
            for r in range(self.world):
                peer = C.create_string_buffer(bytes(handles[r].cpu().numpy().tobytes()), 64)
                _lib.check(L.ptb_exchange_attach_peer(ctx, r, peer))
        else:
            _lib.check(L.ptb_exchange_init_format(ctx, slots, fmt))
            handle = torch.zeros(64, dtype=torch.uint8, device=self.device)
            if self.rank == 0:
                raw = C.create_string_buffer(64)
                _lib.check(L.ptb_exchange_handle(ctx, raw))
                handle.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
            dist.broadcast(handle, src=0)
            if self.rank != 0:
                raw = C.create_string_buffer(bytes(handle.cpu().numpy().tobytes()), 64)
                _lib.check(L.ptb_exchange_attach(ctx, raw))
        dist.barrier()

    def _consume_pending(self, consumer=None):
        """Acquire, hand to `consumer`, and release every rendered frame this rank is the root of.  Returns the last one."""
        import ctypes as C

        import torch

        from . import _lib

        L, ctx = self.tracer._L, self.tracer._ctx
        full = None
        for _ in range(_lib.check(L.ptb_exchange_pending(ctx))):
            ptr = C.c_void_p()
            _lib.check(L.ptb_exchange_acquire(ctx, C.byref(ptr)))
            full = self._slot_tensors.get(ptr.value)
            if full is None:           # wrapping a raw pointer costs tens of microseconds: once per slot, not once per frame
                full = torch.as_tensor(_DeviceBuffer(ptr.value, (self.height, self.width, self.channels)), device=self.device)
                self._slot_tensors[ptr.value] = full
            if consumer is not None:
                consumer(full)
            _lib.check(L.ptb_exchange_release(ctx))
        if full is not None:
            self._last_full = full
        return full

    def last_root(self) -> int:
        """The rank that holds the assembled image of the frame rendered last."""
        from . import _lib
        return _lib.check(self.tracer._L.ptb_exchange_root(self.tracer._ctx, -1)) if self.fused else 0

    def step_fused(self, consumer=None):
        """One frame through the fused path: every rank renders (trace + blend-and-scatter kernels); the frame's root then
        waits, in stream order, until all ranks' pixels have landed, lets `consumer(full_image_tensor)` enqueue its work on the
        current stream, and releases the slot.  Returns the full-image tensor on the root (valid until `slots` frames later)."""
        self.tracer.Render()
        return self._consume_pending(consumer)

    def step_batch(self, frames: int, consumer=None):
        """`frames` frames through the fused path with frame batching (PathTracer.SetBatch): every rank traces them with one
        megakernel launch per batch; one blend-and-scatter kernel per batch still delivers every frame's pixels to its root, so
        a root acquires / consumes / releases one slot per frame exactly as step_fused does."""
        if not self.fused:
            raise RuntimeError("step_batch needs the fused exchange (per-frame NCCL gathers cannot be batched)")
        full = None
        left = int(frames)
        roots = self.world if self.rotate else 1
        while left > 0:
            # a chunk never exceeds the slot ring(s): its blends may only wait for releases that are already enqueued
            chunk = min(left, self.slots * roots)
            left -= chunk
            self.tracer.Render(chunk)
            got = self._consume_pending(consumer)
            full = got if got is not None else full
        return full

    def exchange_ok(self) -> None:
        from . import _lib
        _lib.check(self.tracer._L.ptb_exchange_status(self.tracer._ctx))

    def render(self, frames: int = 1) -> None:
        """Render `frames` frames.  With the fused exchange every frame occupies a slot on rank 0 until it is released, so the
        frames go through step_batch, which acquires and releases in chunks no longer than the slot ring (a plain
        PathTracer.Render(frames) with frames > slots is refused by the library: it would wait for its own releases)."""
        if self.fused:
            self.step_batch(frames)
        else:
            self.tracer.Render(frames)

    def _start_gather(self):
        """Snapshot the local stripes (the next frame overwrites them in place) and start the one collective of the frame."""
        import torch.distributed as dist

        k = self._k
        self._k ^= 1
        self.staging[k].copy_(self.local, non_blocking=True)
        if self.world == 1:
            work = None
            if self.rank == 0:
                self.gathered[k][0].copy_(self.staging[k], non_blocking=True)
        elif self.rank == 0:
            work = dist.gather(self.staging[k], list(self.gathered[k].unbind(0)), dst=0, async_op=True)
        else:
            work = dist.gather(self.staging[k], None, dst=0, async_op=True)
        self._pending = (k, work)

    def _finish_gather(self):
        """Wait (in stream order) for the outstanding gather; rank 0 de-interleaves it into a row-major image."""
        import ctypes as C

        from . import _lib

        if self._pending is None:
            return self._last_full
        k, work = self._pending
        self._pending = None
        if work is not None:
            work.wait()
        if self.rank == 0:
            _lib.check(self.tracer._L.ptb_deinterleave_device(self.tracer._ctx, C.c_void_p(self.gathered[k].data_ptr()),
                                                              C.c_void_p(self.full[k].data_ptr())))
            self._last_full = self.full[k]
        return self._last_full

    def gather(self):
        self._start_gather()
        return self._finish_gather()

    def step(self):
        """Render the next frame, complete the PREVIOUS frame's gather (which ran beside this render on NCCL's stream), then
        start this frame's gather.  Returns the previous frame's full image on rank 0 (None for the first step)."""
        if self.fused:
            return self.step_fused()
        self.tracer.Render()
        done = self._finish_gather() if self._pending is not None else None
        self._start_gather()
        return done

    def flush(self):
        if self.fused:
            return self._last_full
        return self._finish_gather()


def gpu_numa_cpus(device_index: int):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None when that cannot be told (no NVML, no sysfs, single node)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        return cpus or None
    except Exception:      # noqa: BLE001
        return None


class SharedHostFrame:
    """`buffers` full-frame host images in ONE pinned mapping shared by every rank of the node (a /dev/shm file, mapped and
    cudaHostRegister'ed by each process), the destination of PathTracer.ReadResultScatterAsync: each GPU writes its stripes
    over its own PCIe link and the frame is complete on the host without a device-side gather.

    Collective: every rank of the default process group must construct it together.  Raises on any rank => raises on all."""

    def __init__(self, frame_bytes: int, buffers: int, rank: int, world: int, register: bool = True, stripe_bytes: int = 0,
                 numa_cpus=None):
        """register=False maps the frame without pinning it for CUDA (CPU-only tests of the collective set-up).
        stripe_bytes > 0: NUMA placement — before the buffer is pinned, every rank touches the pages of its own stripes
        (stripe s of `stripe_bytes` bytes belongs to rank s % world) while running on `numa_cpus` (the CPUs next to its GPU,
        see gpu_numa_cpus), so the pages a GPU writes over PCIe live in the memory of its own socket.  Without it the whole
        frame sits on whichever node pinned it first and half the GPUs write across the socket interconnect."""
        import os
        import secrets

        import torch
        import torch.distributed as dist

        self.frame_bytes, self.buffers = int(frame_bytes), int(buffers)
        self.nbytes = ((self.frame_bytes + 4095) // 4096 * 4096) * self.buffers
        self.stride = self.nbytes // self.buffers
        self._registered = False
        self.tensor = None
        name = [None]
        ok = 1
        err = ""
        try:
            if rank == 0:
                st = os.statvfs("/dev/shm")
                if st.f_bavail * st.f_frsize < self.nbytes + (16 << 20):
                    raise OSError(f"/dev/shm has {st.f_bavail * st.f_frsize} bytes free, need {self.nbytes}")
                name[0] = f"/dev/shm/ptb200_frame_{os.getpid()}_{secrets.token_hex(4)}"
                with open(name[0], "wb") as f:
                    f.truncate(self.nbytes)
        except Exception as exc:      # noqa: BLE001
            ok, err = 0, str(exc)
        if world > 1:
            dist.broadcast_object_list(name, src=0)
        self.path = name[0]
        try:
            if ok and self.path:
                self.tensor = torch.from_file(self.path, shared=True, size=self.nbytes, dtype=torch.uint8)
                self.numa = self._first_touch(rank, world, int(stripe_bytes), numa_cpus) if stripe_bytes > 0 else "off"
                if world > 1 and stripe_bytes > 0:
                    dist.barrier()         # every page has its owner before anybody pins the whole mapping
                if register:
                    rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), self.nbytes, 1)     # 1 = cudaHostRegisterPortable
                    if int(rc) != 0:
                        raise RuntimeError(f"cudaHostRegister failed: {rc}")
                    self._registered = True
            else:
                ok = 0
        except Exception as exc:      # noqa: BLE001
            ok, err = 0, str(exc)
        if world > 1:
            flag = torch.tensor([ok], dtype=torch.int32, device="cuda" if register else "cpu")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            all_ok = int(flag.item())
        else:
            all_ok = ok
        if rank == 0 and self.path and os.path.exists(self.path):
            os.unlink(self.path)               # every rank has it mapped (or has given up): the name can go
        if world > 1:
            # second reduction = a barrier that works on both backends: when the constructor returns, the name is gone everywhere
            done = torch.zeros(1, dtype=torch.int32, device="cuda" if register else "cpu")
            dist.all_reduce(done)
        if not all_ok:
            self.close()
            raise RuntimeError(f"shared host frame unavailable on some rank ({err or 'see the other ranks'})")

    def _first_touch(self, rank: int, world: int, stripe_bytes: int, numa_cpus) -> str:
        """Writes one byte into every page of this rank's stripes (all buffers), pinned to `numa_cpus` while doing so."""
        import os
        old = None
        note = "first touch without CPU binding"
        try:
            if numa_cpus:
                old = os.sched_getaffinity(0)
                allowed = set(numa_cpus) & set(old)
                if allowed:
                    os.sched_setaffinity(0, allowed)
                    note = f"first touch on CPUs {min(allowed)}-{max(allowed)}"
            view = self.tensor.numpy()
            page = 4096
            for b in range(self.buffers):
                base = b * self.stride
                s = rank
                while s * stripe_bytes < self.frame_bytes:
                    lo = base + s * stripe_bytes
                    hi = min(base + (s + 1) * stripe_bytes, base + self.frame_bytes)
                    first = (lo + page - 1) // page * page if s > 0 else lo // page * page     # a page shared with the previous stripe is its owner's
                    view[first:hi:page] = 0
                    s += world
        finally:
            if old is not None:
                os.sched_setaffinity(0, old)
        return note

    def ptr(self, k: int) -> int:
        return self.tensor.data_ptr() + (k % self.buffers) * self.stride

    def view(self, k: int, shape, dtype):
        import torch
        n = 1
        for d in shape:
            n *= d
        flat = self.tensor[(k % self.buffers) * self.stride:(k % self.buffers) * self.stride + n * torch.empty((), dtype=dtype).element_size()]
        return flat.view(dtype).view(*shape)

    def close(self) -> None:
        import torch
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self._registered = False
        self.tensor = None
