"""A/B of the BVH threshold on the default scene (55 primitives) and a mid-size scene: ms/frame with the brute-force fold
vs the shared-memory BVH, plus a bitwise check between the two.  Run under gpurun; numbers are for A/B decisions only."""
import os, sys, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200
sc = ptb200.scene
cam = sc.default_camera()
for name, scn, depth in (("default 48+7", sc.load_default_scene(), 13), ("synthetic 96+24", sc.synthetic_scene(96, 24, seed=5), 8),
                         ("synthetic 32+8", sc.synthetic_scene(32, 8, seed=6), 8)):
    crcs = []
    for thr in (100000, 16):
        p = ptb200.PathTracer(None, 1920, 1080, depth, 1, 20.0, 0.14, max_spheres=scn.max_spheres, max_cuboids=scn.max_cuboids)
        p.SetBvhThreshold(thr)
        p.GenerateAtmosphere(256, 50, 15, 0.5, 15.0); p.LoadScene(scn); p.SetCamera(cam)
        p.Render(3); p.Synchronize()
        best = 1e9
        for _ in range(3):
            p.ResetRenderer(); p.Render(20); best = min(best, p.LastRenderMs() / 20)
        p.ResetRenderer(); p.Render(2); a = p.Result
        crcs.append(zlib.crc32(a.tobytes()))
        print(f"{name}: threshold {thr}: nodes {p.BvhNodes} always-tested {p.SceneInfo(1)}: {best:.4f} ms/frame -> {1920*1080/best/1e3:.0f} Msamples/s crc={crcs[-1]:08x}", flush=True)
        p.Dispose()
    print(f"{name}: BVH == brute force bitwise: {crcs[0] == crcs[1]}", flush=True)
