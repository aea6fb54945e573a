"""Per-frame floor: host enqueue cost of Render() (ctypes + CUDA API calls) and GPU-side fixed cost, on a tiny image."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200
sc = ptb200.scene
cam, scene = sc.default_camera(), sc.load_default_scene()
for ov in (1, 2):
    p = ptb200.PathTracer(None, 32, 32, 13, 1, 20.0, 0.14)
    p.GenerateAtmosphere(64, 10, 4, 0.5, 15.0); p.LoadScene(scene); p.SetCamera(cam); p.SetOverlap(ov)
    p.Render(50); p.Synchronize()
    n = 3000
    t0 = time.perf_counter()
    for _ in range(n):
        p.Render()
    t_enq = time.perf_counter() - t0
    p.Synchronize()
    t_all = time.perf_counter() - t0
    t0 = time.perf_counter(); p.Render(n); t_enq2 = time.perf_counter() - t0; p.Synchronize(); t_all2 = time.perf_counter() - t0
    print(f"overlap {ov}: python Render() x{n}: enqueue {t_enq/n*1e6:.1f} us/frame, total {t_all/n*1e6:.1f}; one ptb_render_frames({n}): enqueue {t_enq2/n*1e6:.1f}, total {t_all2/n*1e6:.1f} us/frame", flush=True)
    p.Dispose()
