"""Exploratory parity + timing report (run under gpurun).  Prints mismatch statistics CUDA-vs-oracle; the pass/fail
versions of these checks live in tests/test_parity_gpu.py."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200  # noqa: E402
from oracle import oracle as O  # noqa: E402

sc = ptb200.scene
out = {}


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def same(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))


t0 = time.time()
env = O.atmosphere(256, sc.atmosphere_ubo_bytes(), sc.atmosphere_light_pos(0.5), 15.0, 50, 15)
print("oracle atmosphere s", time.time() - t0)
scene = sc.load_default_scene()
cam = sc.default_camera()

pt = ptb200.PathTracer(env, 256, 256, 13, 1, 20.0, 0.14)
pt.LoadScene(scene)
pt.SetCamera(cam)

# ---- unit probes
x = np.linspace(0, 6.2831855, 1000003, dtype=np.float32)
s, c = O.sincos(x)
d = pt.DebugEval(0, x, x.size, 2 * x.size).reshape(-1, 2)
print("sincos mismatches", int((~same(d[:, 0], s)).sum()), int((~same(d[:, 1], c)).sum()))
x = np.concatenate([np.linspace(-110, 90, 1000003, dtype=np.float32), np.array([np.nan, np.inf, -np.inf, 0.0, -0.0], np.float32)])
print("exp mismatches", int((~same(pt.DebugEval(1, x, x.size, x.size), O.exp(x))).sum()))
rng = np.random.default_rng(1)
ab = rng.standard_normal((200000, 2)).astype(np.float32) * np.float32(10) ** rng.integers(-30, 30, (200000, 2)).astype(np.float32)
ab[:8] = np.array([[0.0, -0.0], [-0.0, 0.0], [np.nan, 1], [1, np.nan], [np.inf, 1], [0.0, 4.0], [-0.0, 0.0], [np.nan, np.nan]], np.float32)
r = pt.DebugEval(5, ab, ab.shape[0], 4 * ab.shape[0]).reshape(-1, 4)
print("fmin(+0,-0),( -0,+0) bits:", [hex(int(v)) for v in bits(r[:2, :2]).ravel()])
with np.errstate(all="ignore"):
    print("rcp mismatches", int((~same(r[:, 2], np.float32(1) / ab[:, 0])).sum()), "sqrt mismatches", int((~same(r[:, 3], np.sqrt(ab[:, 1]))).sum()))
dirs = rng.standard_normal((300000, 3)).astype(np.float32)
dirs[:6] = np.eye(3, dtype=np.float32).repeat(2, 0) * np.array([1, -1] * 3, np.float32)[:, None]
dirs[6:14] = np.array([[sx, sy, sz] for sx in (1, -1) for sy in (1, -1) for sz in (1, -1)], np.float32)
dirs[14] = np.nan
dirs[15] = 0
print("env lookup mismatches", int((~same(pt.DebugEval(3, dirs, dirs.shape[0], 3 * dirs.shape[0]).reshape(-1, 3), O.texture_cube(env, dirs))).any(axis=1).sum()))
o = (rng.random((200000, 3)).astype(np.float32) - np.float32(0.5)) * np.array([40, 25, 25], np.float32) + np.array([0, 0, -10], np.float32)
dd = rng.standard_normal((200000, 3)).astype(np.float32)
dd /= np.linalg.norm(dd, axis=1, keepdims=True).astype(np.float32)
rays = np.concatenate([o, dd], axis=1).astype(np.float32)
ubo = scene.ubo_bytes()
ref = O.ray_trace(rays, ubo, 256, 48, 7)
for op in (4, 6):
    got = pt.DebugEval(op, rays, rays.shape[0], 12 * rays.shape[0]).reshape(-1, 12)
    print("trace op", op, "mismatching rays", int((~same(got, ref)).any(axis=1).sum()), "hits", int(ref[:, 0].sum()), "inside", int(ref[:, 2].sum()))


# ---- image parity
def run_case(name, W, H, frames, spp=1, depth=13, focal=20.0, ap=0.14, scn=scene, kernel=0, crop=None):
    basic = sc.basic_data_bytes(cam, W, H)
    ubo = scn.ubo_bytes()
    p = ptb200.PathTracer(env, W, H, depth, spp, focal, ap, max_spheres=scn.max_spheres, max_cuboids=scn.max_cuboids)
    p.LoadScene(scn)
    p.SetCamera(cam)
    p.SetKernel(kernel)
    ref = np.zeros((H, W, 4), np.float32)
    worst = 0
    for f in range(frames):
        p.Render()
        O.render(ref, basic, ubo, env, frame=f, spp=spp, ray_depth=depth, focal_length=focal, aperture_diameter=ap,
                 n_spheres=len(scn.spheres), n_cuboids=len(scn.cuboids), max_spheres=scn.max_spheres, rows=crop)
        got = p.Result
        if crop:
            bad = ~same(got[crop[0]:crop[1]], ref[crop[0]:crop[1]]).all(axis=2)
        else:
            bad = ~same(got, ref).all(axis=2)
        worst = max(worst, int(bad.sum()))
        if bad.any() and f < 2:
            ys, xs = np.nonzero(bad)
            print("   first bad px", ys[0] + (crop[0] if crop else 0), xs[0], got[ys[0] + (crop[0] if crop else 0), xs[0]], ref[ys[0] + (crop[0] if crop else 0), xs[0]])
    print(f"case {name}: {W}x{H} frames={frames} spp={spp} kernel={kernel} max mismatching pixels/frame = {worst}; nonfinite px = {int((~np.isfinite(got)).any(axis=2).sum())}")
    out[name] = worst
    p.Dispose()


run_case("C1", 256, 256, 8)
run_case("C1-naive", 256, 256, 4, kernel=1)
run_case("ragged", 251, 123, 3)
run_case("spp4", 128, 128, 3, spp=4)
run_case("depth1", 128, 128, 2, depth=1)
run_case("dof", 160, 120, 2, focal=5.0, ap=0.5)
run_case("pinhole", 160, 120, 2, focal=20.0, ap=0.0)
run_case("1080p-crop", 1920, 1080, 2, crop=(500, 532))
syn = sc.synthetic_scene(1024, 256)
run_case("C3-small", 192, 108, 2, depth=8, scn=syn)
empty = sc.Scene()
run_case("empty", 64, 64, 2, scn=empty)

# ---- timing, both kernels, 1080p
for kernel in (0, 1):
    p = ptb200.PathTracer(env, 1920, 1080, 13, 1, 20.0, 0.14)
    p.LoadScene(scene); p.SetCamera(cam); p.SetKernel(kernel)
    p.Render(5); p.Synchronize()
    p.Render(50)
    ms = p.LastRenderMs() / 50
    print(f"kernel {kernel}: {ms:.3f} ms/frame  -> {1920*1080/ms/1e3:.1f} Msamples/s")
    out[f"ms_kernel{kernel}"] = ms
    p.SetStats(True); p.Render(1); st = p.ReadStats(); p.SetStats(False)
    print("  stats", st, "bounces/sample", st["bounces"] / max(1, st["samples"]))
    p.Dispose()
p = ptb200.PathTracer(env, 1920, 1080, 8, 1, 20.0, 0.14, max_spheres=1024, max_cuboids=256)
p.LoadScene(syn); p.SetCamera(cam)
p.Render(2); p.Synchronize(); p.Render(5)
ms = p.LastRenderMs() / 5
print(f"C3 mega: {ms:.3f} ms/frame -> {1920*1080/ms/1e3:.1f} Msamples/s")
out["ms_c3"] = ms
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)
