"""Quick kernel iteration loop (run under gpurun): ms/frame of the megakernel on C2 / C3 + a bitwise check against the
proxy kernel (which tests/ hold to the oracle).  Not a bench: numbers here are for A/B decisions only."""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200  # noqa: E402

sc = ptb200.scene
cam = sc.default_camera()
scene = sc.load_default_scene()


def run(name, scn, W, H, depth, spp=1, frames=30, check=True, ms=None):
    p = ptb200.PathTracer(None, W, H, depth, spp, 20.0, 0.14, max_spheres=scn.max_spheres, max_cuboids=scn.max_cuboids)
    p.GenerateAtmosphere(256, 50, 15, 0.5, 15.0)
    p.LoadScene(scn); p.SetCamera(cam)
    p.Render(3); p.Synchronize()
    best = 1e9
    for _ in range(3):
        p.ResetRenderer(); p.Render(frames); best = min(best, p.LastRenderMs() / frames)
    line = f"{name}: {best:.4f} ms/frame -> {W*H*spp/best/1e3:.1f} Msamples/s"
    if check:
        p.ResetRenderer(); p.Render(2); a = p.Result
        p.SetKernel(1); p.ResetRenderer(); p.Render(2); b = p.Result
        same = (a.view(np.uint32) == b.view(np.uint32)).all()
        line += f" | mega==proxy: {bool(same)} crc={zlib.crc32(a.tobytes()):08x}"
    print(line, flush=True)
    p.Dispose()


run("C2 default 1080p", scene, 1920, 1080, 13)
run("C2 spp4", scene, 1920, 1080, 13, spp=4, frames=8)
run("C1 256^2", scene, 256, 256, 13, frames=100)
run("C2/8 (8-GPU share) 1920x135", scene, 1920, 135, 13, frames=100)
syn = sc.synthetic_scene(1024, 256)
run("C3 synthetic 1080p", syn, 1920, 1080, 8, frames=4)
