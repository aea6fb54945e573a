"""Single-GPU timing of every BASELINE config (C1..C5) with the product kernel; writes gpurun_out/configs_<round>.json + .md
(round = argv[1], default r02).  Batched frames (ptb_set_batch default), CUDA events around 3 runs, best run reported, for
the exact and the fast build; every case is first checked bit for bit (exact build) against the proxy kernel (the proxy is
held to the CPU oracle by tests/)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200  # noqa: E402

sc = ptb200.scene
cam = sc.default_camera()
default = sc.load_default_scene()
rows = []


def run(name, scene, W, H, depth, spp=1, focal=20.0, ap=0.14, frames=30):
    p = ptb200.PathTracer(None, W, H, depth, spp, focal, ap, max_spheres=scene.max_spheres, max_cuboids=scene.max_cuboids)
    p.GenerateAtmosphere(256, 50, 15, 0.5, 15.0); p.LoadScene(scene); p.SetCamera(cam)
    p.Render(2); a = p.Result
    p.SetKernel(1); p.ResetRenderer(); p.Render(2); b = p.Result; p.SetKernel(0)
    same = bool((a.view(np.uint32) == b.view(np.uint32)).all())
    p.SetStats(True); p.ResetRenderer(); p.Render(1); st = p.ReadStats(); p.SetStats(False)
    ms = {}
    for prec_name, prec in (("exact", ptb200.PRECISION_EXACT), ("fast", ptb200.PRECISION_FAST)):
        p.SetPrecision(prec)
        p.ResetRenderer(); p.Render(frames); p.Synchronize()
        best = 1e9
        for _ in range(3):
            p.ResetRenderer(); p.Render(frames); best = min(best, p.LastRenderMs() / frames)
        ms[prec_name] = best
    row = dict(config=name, width=W, height=H, ray_depth=depth, spp=spp, focal=focal, aperture=ap, ms_per_frame=ms["exact"], ms_per_frame_fast=ms["fast"],
               msamples_per_s=W * H * spp / ms["exact"] / 1e3, msamples_per_s_fast=W * H * spp / ms["fast"] / 1e3, fold={0: "brute force", 1: "BVH", 2: "ray-classification table", 3: "grid"}[p.SceneInfo(4)],
               bounces_per_sample=st["bounces"] / max(1, st["samples"]), bitwise_equals_proxy=same)
    rows.append(row)
    print(row, flush=True)
    p.Dispose()


run("C1 default 256x256", default, 256, 256, 13, frames=512)
run("C2 default 1920x1080", default, 1920, 1080, 13, frames=128)
run("C3 synthetic 1024 spheres + 256 cuboids 1920x1080 rayDepth 8", sc.synthetic_scene(1024, 256), 1920, 1080, 8, frames=32)
run("C4 default 3840x2160 on ONE GPU (the 8-GPU run is in r02_multigpu.md)", default, 3840, 2160, 13, frames=48)
for ap in (0.0, 0.05, 0.14, 0.3, 0.5):
    for focal in (1.0, 5.0, 20.0, 50.0):
        run(f"C5 DoF aperture {ap} focal {focal}", default, 1920, 1080, 13, focal=focal, ap=ap, frames=64)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
ROUND = sys.argv[1] if len(sys.argv) > 1 else "r02"
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"configs_{ROUND}.json"), "w"), indent=1)
with open(os.path.join(ROOT, "gpurun_out", f"configs_{ROUND}.md"), "w") as f:
    f.write("| config | fold | exact ms/frame | exact Msamples/s | fast ms/frame | fast Msamples/s | bounces/sample | exact == proxy bitwise |\n|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        f.write(f"| {r['config']} | {r['fold']} | {r['ms_per_frame']:.4f} | {r['msamples_per_s']:.0f} | {r['ms_per_frame_fast']:.4f} | {r['msamples_per_s_fast']:.0f} | "
                f"{r['bounces_per_sample']:.2f} | {r['bitwise_equals_proxy']} |\n")
