"""Latency of a geometry edit (what a GUI drag costs before the next frame): scene repack + acceleration-structure rebuild.
Default scene: ray-classification table rebuild on the GPU (fp64 kernel); C3: grid rebuild on the host + upload."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200
sc = ptb200.scene


def probe(name, scene, depth, setup=None):
    p = ptb200.PathTracer(None, 1920, 1080, depth, 1, 20.0, 0.14, max_spheres=scene.max_spheres, max_cuboids=scene.max_cuboids)
    p.GenerateAtmosphere(64, 10, 4, 0.5, 15.0)
    if setup:
        setup(p)
    p.LoadScene(scene); p.SetCamera(sc.default_camera()); p.Render(2); p.Synchronize()
    ts = []
    for k in range(5):
        s0 = scene.spheres[0]
        moved = np.array(s0.GetGPUFriendlyData(), dtype=np.float32).copy() if hasattr(s0, "GetGPUFriendlyData") else None
        t0 = time.perf_counter()
        if moved is not None:
            moved.reshape(-1)[0] += 0.01 * (k + 1)                       # drag the first sphere along x
            p.GameObjectsUBO.SubData(0, moved.nbytes, moved)
        else:
            p.SetRayClassification(1, 18, 16)
        p.SceneInfo(4)                                                   # forces the rebuild (as the next Render() would)
        p.Synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{name}: geometry edit -> structures rebuilt in {min(ts):.2f} ms (median {sorted(ts)[2]:.2f}); fold {p.SceneInfo(4)}", flush=True)
    p.Dispose()


probe("default scene, table 18x16 (32.5 MiB)", sc.load_default_scene(), 13)
probe("default scene, table 24x16 (54 MiB)", sc.load_default_scene(), 13, setup=lambda p: p.SetRayClassification(1, 24, 16))
probe("C3 1280 primitives, grid (host build)", sc.synthetic_scene(1024, 256), 8)
probe("C3 1280 primitives, BVH (host build)", sc.synthetic_scene(1024, 256), 8, setup=lambda p: p.SetLargeSceneMode(0))
