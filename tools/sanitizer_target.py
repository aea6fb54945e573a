"""Small run over this round's kernels for compute-sanitizer (tools/r02_sanitizer.sh): ray-classification fold + completion
queue + batch blend (exact and fast), grid and BVH folds, SPP > 1 (ring without the queue), the three read-back formats,
the fast atmosphere kernel."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200  # noqa: E402

sc = ptb200.scene
cam = sc.default_camera()
W, H = 128, 72


def tracer(scene, depth, spp=1):
    p = ptb200.PathTracer(None, W, H, depth, spp, 20.0, 0.14, max_spheres=scene.max_spheres, max_cuboids=scene.max_cuboids)
    p.GenerateAtmosphere(32, 10, 4, 0.5, 15.0); p.LoadScene(scene); p.SetCamera(cam)
    return p


p = tracer(sc.load_default_scene(), 13)
for prec in (ptb200.PRECISION_EXACT, ptb200.PRECISION_FAST):
    p.SetPrecision(prec); p.ResetRenderer(); p.Render(35); p.Synchronize()          # two batches of 16 + three single frames
    print("default scene, precision", prec, "fold", p.SceneInfo(4), "mean", float(p.Result[..., :3].mean()), flush=True)
host = np.empty((H, W, 4), np.float32)
for fmt in (ptb200.FORMAT_RGBA32F, ptb200.FORMAT_RGB32F, ptb200.FORMAT_RGBA8):
    p.ReadResultAsync(host.ctypes.data, fmt)
p.Synchronize()
p.GenerateAtmosphere(32, 10, 4, 0.3, 15.0, fast=True); p.Render(2); p.Synchronize()
p.Dispose()

p = tracer(sc.load_default_scene(), 13, spp=2)
p.Render(3); p.Synchronize(); print("SPP 2 mean", float(p.Result[..., :3].mean()), flush=True)
p.Dispose()

big = sc.synthetic_scene(160, 40)
for mode in (1, 0):
    p = tracer(big, 8)
    p.SetLargeSceneMode(mode)
    for prec in (ptb200.PRECISION_EXACT, ptb200.PRECISION_FAST):
        p.SetPrecision(prec); p.ResetRenderer(); p.Render(18); p.Synchronize()
        print("200 primitives, large-scene mode", mode, "precision", prec, "fold", p.SceneInfo(4), "mean", float(p.Result[..., :3].mean()), flush=True)
    p.Dispose()
print("done", flush=True)
