"""Builds kernel variants (-D defines) into variants/*.so for A/B runs: PTB_LIB=variants/x.so python tools/perf_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from importlib import import_module  # noqa: E402

B = import_module("opentk-pathtracer_b200._build")
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    out = os.path.join(ROOT, "variants", f"{name}.so")
    B.build(force=True, defines=[d for d in defs.split(",") if d], out=out)
    print("built", out)
