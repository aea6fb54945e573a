"""End-to-end (render + pipelined read-back to pinned host memory) per overlap mode.  Run under gpurun."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200
sc = ptb200.scene
cam, scene = sc.default_camera(), sc.load_default_scene()
W, H = 1920, 1080
bufs = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
for ov in (1, 2, 3, 1, 3):
    p = ptb200.PathTracer(None, W, H, 13, 1, 20.0, 0.14)
    p.GenerateAtmosphere(256, 50, 15, 0.5, 15.0); p.LoadScene(scene); p.SetCamera(cam); p.SetOverlap(ov)
    for i in range(5):
        p.Render(); p.ReadResultAsync(bufs[i & 1].data_ptr())
    p.Synchronize()
    n = 200
    t0 = time.perf_counter()
    for i in range(n):
        p.Render(); p.ReadResultAsync(bufs[i & 1].data_ptr())
    p.Synchronize()
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    for i in range(n):
        p.ReadResultAsync(bufs[i & 1].data_ptr())
    p.Synchronize()
    dc = time.perf_counter() - t0
    print(f"overlap {ov}: render+read {dt/n*1e3:.3f} ms/frame ({W*H/(dt/n)/1e6:.0f} Msamples/s); read-back alone {dc/n*1e3:.3f} ms ({W*H*16/(dc/n)/1e9:.1f} GB/s)", flush=True)
    p.Dispose()
