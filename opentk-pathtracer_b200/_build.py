"""Build recipe for libptb200.so (nvcc, sm_100a only, in-tree so the .so travels with the snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# PTB_LIB points the loader at an alternative build (kernel A/B experiments, tools/build_variants.py); never set in tests/bench
LIB = os.environ.get("PTB_LIB") or os.path.join(HERE, "libptb200.so")
DEPS = ["ptb_abi.cu", "ptb_fast.cu", "ptb_fast.h", "ptb_kernels.cuh", "ptb_math.cuh", os.path.join("..", "..", "include", "ptb200.h")]

COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
# Two translation units, two arithmetic contracts:
#   ptb_abi.cu   -fmad=false: the evaluation model forbids implicit contraction (explicit __fmaf_rn only); parity depends on it
#   ptb_fast.cu  -fmad=true -ftz=true: the fast build of the megakernel (ptb_set_precision), MUFU approximations in ptb_math.cuh
UNITS = [("ptb_abi.cu", ["-fmad=false"]), ("ptb_fast.cu", ["-fmad=true", "-ftz=true"])]
LINK = ["-shared", "-cudart", "static"]


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
    """Compiles the two translation units (in parallel) and links them into one shared library."""
    from concurrent.futures import ThreadPoolExecutor
    out = out or LIB
    if force or is_stale() or out != LIB:
        env = dict(os.environ)
        env.pop("CC", None)   # the image exports a CC without OpenMP specs; nvcc should use the system g++
        env.pop("CXX", None)
        objdir = os.path.join(HERE, "build", os.path.splitext(os.path.basename(out))[0])
        os.makedirs(objdir, exist_ok=True)

        def compile_unit(unit):
            src, flags = unit
            obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
            cmd = [nvcc(), *COMMON, *flags, *[f"-D{d}" for d in defines], "-c", "-o", obj, src]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
            res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True, env=env)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
            if verbose:
                print(res.stderr)
            return obj

        with ThreadPoolExecutor(max_workers=len(UNITS)) as pool:
            objs = list(pool.map(compile_unit, UNITS))
        res = subprocess.run([nvcc(), *COMMON, *LINK, "-o", out, *objs], cwd=CSRC, capture_output=True, text=True, env=env)
        if res.returncode != 0:
            raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return out
