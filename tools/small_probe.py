"""Small-tile timing (the 8-GPU share of a 1080p frame) + bitwise check against the proxy.  Run under gpurun."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200
sc = ptb200.scene
cam, scene = sc.default_camera(), sc.load_default_scene()
OVERLAPS = [int(v) for v in os.environ.get("PTB_OVERLAPS", "2").split(",")]
GRID_DIV = int(os.environ.get("PTB_GRID_DIV", "1"))
BATCH = int(os.environ.get("PTB_BATCH", "1"))
for (W, H), ov in [((w, h), o) for (w, h) in [(256, 256), (1920, 135), (1920, 270), (1920, 1080)] for o in OVERLAPS]:
    p = ptb200.PathTracer(None, W, H, 13, 1, 20.0, 0.14)
    p.SetOverlap(ov)
    p.SetGridDivisor(GRID_DIV)
    p.SetBatch(BATCH)
    p.GenerateAtmosphere(256, 50, 15, 0.5, 15.0); p.LoadScene(scene); p.SetCamera(cam)
    p.Render(5); p.Synchronize()
    best = 1e9
    for _ in range(3):
        p.ResetRenderer(); p.Render(100); best = min(best, p.LastRenderMs() / 100)
    p.ResetRenderer(); p.Render(3); a = p.Result
    p.SetKernel(1); p.ResetRenderer(); p.Render(3); b = p.Result
    print(f"{W}x{H} overlap {ov} grid/{GRID_DIV} batch {BATCH}: {best*1e3:8.1f} us/frame  bitwise==proxy: {bool((a.view(np.uint32) == b.view(np.uint32)).all())}", flush=True)
    p.Dispose()
