"""Do two persistent launches from independent streams overlap their tails?  Two contexts on one GPU, frames enqueued
alternately without syncs, vs one context doing all frames.  Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200
sc = ptb200.scene
cam, scene = sc.default_camera(), sc.load_default_scene()


def mk(W, H):
    p = ptb200.PathTracer(None, W, H, 13, 1, 20.0, 0.14)
    p.GenerateAtmosphere(256, 50, 15, 0.5, 15.0); p.LoadScene(scene); p.SetCamera(cam)
    p.Render(5); p.Synchronize()
    return p


for (W, H) in [(1920, 135), (1920, 270), (1920, 1080)]:
    a, b = mk(W, H), mk(W, H)
    n = 200
    t0 = time.perf_counter(); a.Render(2 * n); a.Synchronize(); t1 = time.perf_counter() - t0
    a.Synchronize(); b.Synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        a.Render(); b.Render()
    a.Synchronize(); b.Synchronize()
    t2 = time.perf_counter() - t0
    print(f"{W}x{H}: one context {t1/(2*n)*1e6:7.1f} us/frame; two interleaved contexts {t2/(2*n)*1e6:7.1f} us/frame", flush=True)
    a.Dispose(); b.Dispose()
