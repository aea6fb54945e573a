"""The nvcc build of the reference's compute shader (oracle/build_ref.py --cuda: the GL-compute proxy made from the
reference's own source) cross-compiles for sm_100a without a GPU and exports its C entry points.  Running it needs a GPU
(tools/gl_proxy_probe.py, bench.py's `gl_proxy`)."""
import os
import re
import shutil
import subprocess

import pytest

REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REFERENCE) or not (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")),
                    reason="needs /root/reference and nvcc")
@pytest.mark.parametrize("fast", [False, True], ids=["model-exact", "fast-math"])
def test_reference_shader_cross_compiles_for_sm100a(fast):
    from oracle import build_ref
    lib = build_ref.build_cuda(REFERENCE, fast=fast)
    assert os.path.exists(lib)
    syms = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (glref_cuda_[a-z_]+)", syms))
    assert exported == {"glref_cuda_pt_render", "glref_cuda_pt_capacity", "glref_cuda_last_error"}
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    elf = subprocess.run([cuobjdump, "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in elf and not re.search(r"sm_(?!100a)\d+", elf)
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True).stdout
    assert "glsl_dispatch" in sass
    # the model-exact build must not contract a*b+c on its own: every FFMA comes from an explicit fma() of the model
    # (dot products, mat*vec, mix, the polynomial kernels); the fast build is free to use MUFU approximations
    if fast:
        assert "MUFU.SIN" in sass or "MUFU.COS" in sass or "MUFU.EX2" in sass
    else:
        assert "MUFU.SIN" not in sass and "MUFU.COS" not in sass and "MUFU.EX2" not in sass
