#!/usr/bin/env python
"""The GL-compute proxy built from the reference's own source (run on a GPU box): compute.glsl compiled by nvcc
(oracle/build_ref.py --cuda [--fast]) and dispatched in the reference's launch shape — 8x8 groups, one invocation per pixel,
one dispatch per frame.  Prints one JSON object:

  exact  evaluation model of glsl_model.h (-fmad=false): ms per dispatch, Msamples/s, and whether two accumulated frames
         equal the CPU build of the same shader bit for bit;
  fast   -use_fast_math, MUFU built-ins, FMA contraction (roughly a GL driver's code generation): ms per dispatch, Msamples/s,
         per-channel MSE of its first frame against the exact build.

A measurement aid: bench.py runs it in a subprocess (a crash here cannot take the bench down) and attaches the result as
`gl_proxy`.  Never part of the product path.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=20)
    a = ap.parse_args()
    import ptb200
    from oracle import oracle as O, ref as R, ref_cuda
    sc = ptb200.scene
    W, H = a.width, a.height
    cpu = R if R.available() else O
    env = cpu.atmosphere(256, sc.atmosphere_ubo_bytes(), sc.atmosphere_light_pos(0.5), 15.0, 50, 15)
    scene, cam = sc.load_default_scene(), sc.default_camera()
    basic, ubo = sc.basic_data_bytes(cam, W, H), scene.ubo_bytes()
    kw = dict(spp=1, ray_depth=13, focal_length=20.0, aperture_diameter=0.14, n_spheres=48, n_cuboids=7)
    out = {"config": f"default scene, {W}x{H}, SPP 1, rayDepth 13, 8x8 work groups, one dispatch per frame", "frames_timed": a.frames}
    first = {}
    for name, fast in (("exact", False), ("fast", True)):
        try:
            g = ref_cuda.CudaReference(fast=fast)
            img = np.zeros((H, W, 4), np.float32)
            g.render(img, basic, ubo, env, frame=0, frames=2, **kw)
            one = np.zeros((H, W, 4), np.float32)
            g.render(one, basic, ubo, env, frame=0, frames=1, **kw)
            first[name] = one
            scratch = np.zeros((H, W, 4), np.float32)
            g.render(scratch, basic, ubo, env, frame=0, frames=3, **kw)                       # warm-up
            ms = g.render(scratch, basic, ubo, env, frame=3, frames=a.frames, **kw)
            entry = {"ms_per_dispatch": ms, "msamples_per_s": W * H / ms / 1e3}
            if not fast:
                want = np.zeros((H, W, 4), np.float32)
                for f in range(2):
                    cpu.render(want, basic, ubo, env, frame=f, **kw)
                same = (img.view(np.uint32) == want.view(np.uint32)) | (np.isnan(img) & np.isnan(want))
                entry["bit_exact_vs_cpu_build_of_the_same_shader"] = bool(same.all())
                entry["differing_pixels"] = int((~same).any(axis=-1).sum())
            elif "exact" in first:
                d = one[..., :3].astype(np.float64) - first["exact"][..., :3]
                entry["mse_vs_exact_first_frame"] = float(np.nanmean(d * d))
            out[name] = entry
        except Exception as exc:      # noqa: BLE001
            out[name] = {"error": str(exc)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
