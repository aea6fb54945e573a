"""Timing of the two atmosphere producers (exact: the shader as a kernel; fast: tabulated secondary loop).  Run under gpurun."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200
sc = ptb200.scene
p = ptb200.PathTracer(None, 64, 64, 13, 1, 20.0, 0.14)
for size, fast in ((256, False), (256, True), (1024, True), (2048, True), (2048, False)):
    p.GenerateAtmosphere(size, 50, 15, 0.5, 15.0, fast=fast); p.Synchronize()
    t0 = time.perf_counter()
    n = 3 if size >= 2048 and not fast else 10
    for k in range(n):
        p.GenerateAtmosphere(size, 50, 15, 0.5 + 0.01 * k, 15.0, fast=fast)
    p.Synchronize()
    print(f"atmosphere {size}^2 x 6, 50 x 15 steps, {'fast' if fast else 'exact'}: {(time.perf_counter() - t0) / n * 1e3:.2f} ms per regeneration (incl. allocation + cubemap padding)", flush=True)
p.Dispose()
