"""Build recipe for libptb200.so (nvcc, sm_100a only, in-tree so the .so travels with the snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# PTB_LIB points the loader at an alternative build (kernel A/B experiments, tools/build_variants.py); never set in tests/bench
LIB = os.environ.get("PTB_LIB") or os.path.join(HERE, "libptb200.so")
SOURCES = ["ptb_abi.cu"]
DEPS = ["ptb_abi.cu", "ptb_kernels.cuh", "ptb_math.cuh", os.path.join("..", "..", "include", "ptb200.h")]

# -fmad=false: the evaluation model forbids implicit contraction (explicit __fmaf_rn only); parity depends on it.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if os.environ.get("PTB_LIB"):
        return False
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    out = out or LIB
    if force or is_stale() or out != LIB:
        cmd = [nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", out, *SOURCES]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        env = dict(os.environ)
        env.pop("CC", None)   # the image exports a CC without OpenMP specs; nvcc should use the system g++
        env.pop("CXX", None)
        res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True, env=env)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr)
    return out
