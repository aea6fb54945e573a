"""Is the fast build's difference from the exact build bias, or only Monte-Carlo decorrelation?  

A path is chaotic: one ulp flips a Russian-roulette or lobe decision and the rest of that sample is a different, equally
valid, sample.  So fast-vs-exact at matched seeds is bounded by the difference of two INDEPENDENT estimates of the same
image.  Per config: A = exact, frames 0..N-1; B = fast, same frames; M = exact, frames 0..2N-1; C = 2M - A = the exact mean of
frames N..2N-1 (independent of A).  Prints per-channel MSE(A,B), MSE(A,C) and the mean signed difference B - A (bias)
beside its standard error estimated from C - A.  Writes gpurun_out/precision_noise.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200  # noqa: E402

sc = ptb200.scene
cam = sc.default_camera()
out = []
for name, scene, depth, N in (("c2", sc.load_default_scene(), 13, 1024), ("c3", sc.synthetic_scene(1024, 256), 8, 256)):
    W, H = 1920, 1080
    p = ptb200.PathTracer(None, W, H, depth, 1, 20.0, 0.14, max_spheres=scene.max_spheres, max_cuboids=scene.max_cuboids)
    p.GenerateAtmosphere(256, 50, 15, 0.5, 15.0); p.LoadScene(scene); p.SetCamera(cam)

    def mean(prec, frames):
        p.SetPrecision(prec); p.ResetRenderer(); p.Render(frames)
        return p.Result[..., :3].astype(np.float64)
    A = mean(ptb200.PRECISION_EXACT, N)
    B = mean(ptb200.PRECISION_FAST, N)
    M = mean(ptb200.PRECISION_EXACT, 2 * N)
    Cc = 2.0 * M - A
    ok = np.isfinite(A).all(axis=-1) & np.isfinite(B).all(axis=-1) & np.isfinite(Cc).all(axis=-1)
    row = dict(config=name, frames=N, pixels=int(ok.sum()),
               mse_fast_vs_exact=[float(((B - A)[ok][:, k] ** 2).mean()) for k in range(3)],
               mse_independent_exact=[float(((Cc - A)[ok][:, k] ** 2).mean()) for k in range(3)],
               mean_signed_diff_fast_minus_exact=[float((B - A)[ok][:, k].mean()) for k in range(3)],
               mean_signed_diff_independent=[float((Cc - A)[ok][:, k].mean()) for k in range(3)],
               std_error_of_the_mean_diff=[float((Cc - A)[ok][:, k].std() / np.sqrt(ok.sum())) for k in range(3)],
               image_mean=[float(A[ok][:, k].mean()) for k in range(3)],
               pixels_identical_fast_exact=float((np.abs(B - A).max(axis=-1) == 0)[ok].mean()))
    out.append(row)
    print(json.dumps(row), flush=True)
    p.Dispose()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "precision_noise.json"), "w"), indent=1)
