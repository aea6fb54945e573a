import os, sys, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200
sc = ptb200.scene
cam = sc.default_camera()
syn = sc.synthetic_scene(1024, 256)
p = ptb200.PathTracer(None, 1920, 1080, 8, 1, 20.0, 0.14, max_spheres=1024, max_cuboids=256)
if os.environ.get('PTB_GRID_DENSITY'):
    p.SetGridDensity(float(os.environ['PTB_GRID_DENSITY']))
if os.environ.get('PTB_LARGE') == 'bvh':
    p.SetLargeSceneMode(0)
if os.environ.get('PTB_PRECISION') == 'fast':
    p.SetPrecision(1)
p.GenerateAtmosphere(256, 50, 15, 0.5, 15.0); p.LoadScene(syn); p.SetCamera(cam)
p.Render(2); p.Synchronize()
best = 1e9
for _ in range(3):
    p.ResetRenderer(); p.Render(6); best = min(best, p.LastRenderMs() / 6)
p.ResetRenderer(); p.Render(2); a = p.Result
st = {'bounces': 0, 'samples': 1}
if os.environ.get('PTB_PRECISION') != 'fast':
    p.SetStats(True); p.ResetRenderer(); p.Render(1); st = p.ReadStats(); p.SetStats(False)
print(f"C3 [{os.environ.get('PTB_LARGE', 'grid')}, {os.environ.get('PTB_PRECISION', 'exact')}] fold {p.SceneInfo(4)} cells {p.SceneInfo(6)} items {p.SceneInfo(7)} always {p.SceneInfo(1)} staged {p.SceneInfo(2)} B: {best:.3f} ms/frame -> {1920*1080/best/1e3:.0f} Msamples/s crc={zlib.crc32(a.tobytes()):08x} bounces/sample={st['bounces']/st['samples']:.2f}", flush=True)
