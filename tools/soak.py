"""Soak: thousands of pipelined frames with interleaved async read-backs, scene edits and resets; the pipelined result must
equal the in-place (overlap 1) result bit for bit at the end.  Run under gpurun."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200
sc = ptb200.scene
cam = sc.default_camera()
W, H, N = 640, 360, 6000
bufs = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
res = {}
for ov in (2, 1, 3):
    scene = sc.load_default_scene()
    p = ptb200.PathTracer(None, W, H, 13, 1, 20.0, 0.14)
    p.GenerateAtmosphere(128, 20, 8, 0.5, 15.0); p.LoadScene(scene); p.SetCamera(cam); p.SetOverlap(ov)
    t0 = time.perf_counter()
    for i in range(N):
        p.Render()
        if i % 7 == 0:
            p.ReadResultAsync(bufs[(i // 7) & 1].data_ptr())
        if i == 2000:
            scene.spheres[3].Material.Emissiv = np.array([4.0, 1.0, 0.2], np.float32); scene.spheres[3].Upload(p.GameObjectsUBO)
            p.ResetRenderer()
        if i == 4000:
            p.FocalLength = 12.0; p.ApertureDiameter = 0.3; p.ResetRenderer()
    p.Synchronize()
    dt = time.perf_counter() - t0
    res[ov] = p.Result
    print(f"overlap {ov}: {N} frames in {dt:.2f} s ({dt/N*1e6:.1f} us/frame), samples={p.Samples}, finite={bool(np.isfinite(res[ov]).all())}", flush=True)
    p.Dispose()
same = all((res[2].view(np.uint32) == res[k].view(np.uint32)).all() for k in (1, 3))
print("SOAK_OK" if same else "SOAK_MISMATCH", flush=True)
