"""Multi-GPU correctness + timing of both exchange paths (NCCL gather vs fused peer stores).  Launch under torchrun:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py
Rank 0 compares the exchanged image with a single-GPU render of the same frames, bit for bit."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200  # noqa: E402
from importlib import import_module  # noqa: E402

D = import_module("opentk-pathtracer_b200.distributed")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sc = ptb200.scene
scene, cam = sc.load_default_scene(), sc.default_camera()
W, H = (int(v) for v in os.environ.get("SIZE", "1920x1080").split("x"))
FRAMES = int(os.environ.get('FRAMES', '6'))


def make():
    pt = ptb200.PathTracer(None, W, H, 13, 1, 20.0, 0.14, device=local)
    pt.GenerateAtmosphere(256, 50, 15, 0.5, 15.0)
    pt.LoadScene(scene)
    pt.SetCamera(cam)
    return pt


ref = None
if True:
    single = make()
    refs = []
    for _ in range(FRAMES):
        single.Render()
        refs.append(single.Result)
    ref = refs[-1]
    single.Dispose()

ok = True
for fused, rgb, batched, rotate in ((False, False, False, False), (True, False, False, False), (True, True, True, False), (True, True, True, True)):
    pt = make()
    tp = D.TiledPathTracer(pt, rank, world, 8, device=dev, fused=fused, slots=int(os.environ.get('SLOTS', '2')), rgb=rgb, rotate=rotate)
    snap = torch.empty((H, W, tp.channels), dtype=torch.float32, device=dev)
    full = None
    if batched:        # frames traced by batched launches (slot ring = 2: chunks of two), every frame delivered in RGB32F
        pt.SetBatch(4)
        tp.step_batch(FRAMES, consumer=lambda t: snap.copy_(t))       # runs on the root of each frame
    else:
        for f in range(FRAMES):
            if fused:      # the slot is only ours between acquire and release: copy it out inside the consumer callback
                tp.step_fused(consumer=lambda t: snap.copy_(t))
            else:
                tp.step()
    full = snap if fused else tp.flush()
    torch.cuda.synchronize(dev)
    if fused and tp.rotate:       # rotating roots: the last frame was assembled on rank last_root(); bring it to rank 0 for the comparison
        dist.broadcast(snap, src=tp.last_root())
    mine = D.local_rows_of(rank, world, 8, H)
    loc = pt.Result
    print(f'rank {rank} fused={fused}: LOCAL stripes == single-GPU rows: {bool((loc.view(np.uint32) == ref[mine].view(np.uint32)).all())}', flush=True)
    if fused:
        tp.exchange_ok()
    if rank == 0:
        got = full.cpu().numpy()
        if got.shape[2] == 3:          # RGB32F slots: the colour floats, bit for bit
            got = np.concatenate([got, np.ones_like(got[..., :1])], axis=2)
        same = bool((got.view(np.uint32) == ref.view(np.uint32)).all())
        ok &= same
        if not same:
            time.sleep(1.0)
            got2 = full.cpu().numpy()
            if got2.shape[2] == 3:
                got2 = np.concatenate([got2, np.ones_like(got2[..., :1])], axis=2)
            print('   after 1 s the image is correct:', bool((got2.view(np.uint32) == ref.view(np.uint32)).all()), flush=True)
            bad = (got.view(np.uint32) != ref.view(np.uint32)).any(axis=2)
            rows = np.nonzero(bad.any(axis=1))[0]
            for fi, rf in enumerate(refs):
                print(f'   rows of rank 1 equal the single-GPU image after {fi + 1} frame(s):', bool((got[8:16].view(np.uint32) == rf[8:16].view(np.uint32)).all()), flush=True)
            yy, xx = np.argwhere(bad)[len(np.argwhere(bad)) // 2]
            print('   sample bad pixel', int(yy), int(xx), 'got', got[yy, xx], 'refs per frame count:', [rf[yy, xx, 0] for rf in refs], flush=True)
            print('   mismatching rows:', rows[:24].tolist(), '... count', rows.size, 'of', H, '; bad px', int(bad.sum()), '; got', got[rows[0], 0], 'ref', ref[rows[0], 0], flush=True)
        print(f"[{'fused' if fused else 'nccl '}{' rgb batched' if batched else ''}{' rotating roots' if rotate else ''}] {world} GPUs {W}x{H}: exchanged image == single-GPU render: {same}", flush=True)
    dist.barrier()
    # timing: K pipelined steps
    for _ in range(10):
        tp.step()
    tp.flush(); dist.barrier(); torch.cuda.synchronize(dev)
    K = 300
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if batched:
        tp.step_batch(K)
    else:
        for _ in range(K):
            tp.step()
    tp.flush()
    e1.record()
    dist.barrier(); torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if fused:
        tp.exchange_ok()
    if rank == 0:
        print(f"[{'fused' if fused else 'nccl '}] {ms.item()/K*1e3:.1f} us/step -> {W*H*K/(ms.item()*1e-3)/1e6:.0f} Msamples/s", flush=True)
    dist.barrier()
    pt.Dispose()
if rank == 0:
    print("MULTIGPU_OK" if ok else "MULTIGPU_MISMATCH", flush=True)
dist.destroy_process_group()
