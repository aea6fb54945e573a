"""A/B of the ray-classification table on the default scene at 1080p: off vs a sweep of (cells, buckets).  ms/frame in the
in-place mode (megakernel alone) + a CRC of the image (must not change).  Run under gpurun; for A/B decisions only."""
import os, sys, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200
sc = ptb200.scene
cam, scene = sc.default_camera(), sc.load_default_scene()
configs = [(0, 13, 12)] + [(1, c, g) for c, g in ((8, 8), (13, 8), (13, 12), (13, 16), (18, 12), (18, 16))]
if len(sys.argv) > 1:
    configs = [(0, 13, 12)] + [(1, int(a.split(",")[0]), int(a.split(",")[1])) for a in sys.argv[1:]]
for mode, cells, buckets in configs:
    p = ptb200.PathTracer(None, 1920, 1080, 13, 1, 20.0, 0.14)
    p.SetRayClassification(mode, cells, buckets)
    if os.environ.get('PTB_PRECISION') == 'fast':
        p.SetPrecision(1)
    p.SetOverlap(1)
    p.GenerateAtmosphere(256, 50, 15, 0.5, 15.0); p.LoadScene(scene); p.SetCamera(cam)
    p.Render(5); p.Synchronize()
    best = 1e9
    for _ in range(3):
        p.ResetRenderer(); p.Render(20); best = min(best, p.LastRenderMs() / 20)
    p.ResetRenderer(); p.Render(2); crc = zlib.crc32(p.Result.tobytes())
    print(f"rct mode {mode} cells {cells:2d} buckets {buckets:2d} table {p.SceneInfo(5) / 1024:6.1f} MiB: {best:.4f} ms/frame -> {1920 * 1080 / best / 1e3:.0f} Msamples/s crc={crc:08x}", flush=True)
    p.Dispose()
