"""C2 only (default scene, 1080p): the workload ncu captures in tools/ncu_c2.sh.  PTB_PRECISION=fast selects the fast build."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200
sc = ptb200.scene
p = ptb200.PathTracer(None, 1920, 1080, 13, 1, 20.0, 0.14)
if os.environ.get("PTB_PRECISION") == "fast":
    p.SetPrecision(1)
p.SetOverlap(int(os.environ.get("PTB_OVERLAP", "1")))
p.GenerateAtmosphere(256, 50, 15, 0.5, 15.0); p.LoadScene(sc.load_default_scene()); p.SetCamera(sc.default_camera())
p.Render(12); p.Synchronize()
print(f"{p.LastRenderMs() / 12:.4f} ms/frame")
p.Dispose()
