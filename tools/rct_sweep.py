"""Ray-classification table resolution sweep on the batched fast path (C2 1080p and C4 4K): ms per frame for each
(cells, buckets) given on the command line as cells,buckets pairs.  Best of 3 runs of 128 (C2) / 48 (C4) frames."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200
sc = ptb200.scene
for W, H, frames in ((1920, 1080, 128), (3840, 2160, 48)):
    for arg in sys.argv[1:]:
        cells, buckets = (int(v) for v in arg.split(","))
        p = ptb200.PathTracer(None, W, H, 13, 1, 20.0, 0.14)
        p.SetPrecision(ptb200.PRECISION_FAST)
        p.SetRayClassification(1, cells, buckets)
        p.GenerateAtmosphere(256, 50, 15, 0.5, 15.0); p.LoadScene(sc.load_default_scene()); p.SetCamera(sc.default_camera())
        p.Render(frames); p.Synchronize()
        best = 1e9
        for _ in range(3):
            p.ResetRenderer(); p.Render(frames); best = min(best, p.LastRenderMs() / frames)
        print(f"{W}x{H} cells {cells} buckets {buckets}: {best:.4f} ms/frame, {W * H / best / 1e3:.0f} Msamples/s, table {p.SceneInfo(5) / 1024:.1f} MiB", flush=True)
        p.Dispose()
