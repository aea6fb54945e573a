#!/usr/bin/env python
"""oracle/build_ref.py — TEST INFRASTRUCTURE ONLY.  Recipe that compiles the REFERENCE'S OWN SHADERS for the CPU.

    python oracle/build_ref.py [--reference /root/reference] [--keep]

reads the three GLSL files of the hot path where they lie under /root/reference (nothing is copied into the repo),
applies the lexical rewrites R1..R12 below so that g++ accepts the text as C++ against oracle/glsl_shim.hpp (the GLSL
types, swizzles and built-ins), wraps each in its harness (oracle/ref_harness_*.inc: the job the GL driver and the C# host
do — filling UBO members from std140 bytes, binding images, looping gl_GlobalInvocationID) and links
oracle/_ref/libglsl_ref.so.  The translated text only ever exists in a temporary directory under /tmp (left there with --keep for
inspection); oracle/_ref/ is git-ignored and holds the binary plus a manifest with the SHA-256 of each shader it was
built from.

The rewrites do not touch any expression structure: every arithmetic operation, its operands and its order are the
shader's own.  They are:
  R1  drop the `#version` line.
  R2  floating literals `L` become `Float(Lf)`; R3 the type name `float` becomes `Float` (glsl_shim.hpp: GLSL scalar
      semantics instead of C++'s double promotion / compiler constant folding).
  R4  `layout(...) in;` (work-group size) is dropped; `layout(...)` / `uniform` / `restrict` / `writeonly` qualifiers are
      dropped from global declarations.
  R5  interface blocks `[layout(...)] uniform|in|out Name { ... } inst;` become `struct Name { ... } inst;`.
  R6  `out T x` / `inout T x` parameters become `T& x`.
  R7  swizzles `.xyz .rgb .xy .zw` become member calls `.xyz()` ...; `0.0031308.xxx` becomes `vec3(0.0031308)`.
  R8  `void main()` becomes `void glsl_main()`.
  R9  a struct member declared `Material Material;` becomes `struct Material Material;` (C++ name lookup).
  R10 GLSL array-type syntax `mat4[6] InvView;` becomes `mat4 InvView[6];`.
  R11 constructor calls `vecN(...)` / `Ray(...)` become brace initialisation `vecN{...}`: C++ leaves the evaluation order
      of parenthesised arguments unspecified, GLSL §6.1.1 evaluates left to right, braces guarantee it
      (compute.glsl:113 draws two random numbers inside one constructor).  The script then verifies that no other
      statement contains two RNG-advancing calls whose order C++ would not fix.
  R12 GLSL globals are per invocation: mutable globals (`uint rndSeed;`), fragment inputs / outputs, the built-in variables and
      all functions become members of `struct Invocation` (one object per invocation); types, uniforms and uniform blocks
      stay shared at namespace scope; forward declarations are dropped.  The same text compiles for the GPU (GLSL_FN).
  R13 (only with --capacity S C, for BASELINE config 3, which does not fit the shader as shipped) the two array lengths of
      the GameObjectsUBO block, `Spheres[256]` / `Cuboids[64]` (compute.glsl:68-69), become `Spheres[S]` / `Cuboids[C]`;
      output goes to libglsl_ref_<S>x<C>.so.  Nothing else changes.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libglsl_ref.so")
LIB_ALT = os.path.join(OUT_DIR, "libglsl_ref_alt.so")
MANIFEST = os.path.join(OUT_DIR, "manifest.json")

SHADERS = {
    # key: (path under <reference>/OpenTK-PathTracer/res/shaders, harness include, C++ namespace)
    "pt": ("PathTracing/compute.glsl", "ref_harness_pt.inc", "pt"),
    "atmosphere": ("AtmosphericScattering/compute.glsl", "ref_harness_atmosphere.inc", "atmo"),
    "post": ("PostProcessing/fragment.glsl", "ref_harness_post.inc", "post"),
}
# functions that advance the RNG (directly or through callees) — used only by the R11 order check
IMPURE = ("GetPCGHash", "GetRandomFloat01", "CosineSampleHemisphere", "UniformSampleUnitCircle", "BSDF", "Radiance")

CXX = "/usr/bin/g++"   # like oracle/Makefile: the system compiler (an env-provided g++ may lack libgomp)
# same value-preserving flags as oracle/Makefile; -fwrapv: GLSL integer arithmetic wraps
CXXFLAGS = ["-std=c++17", "-O3", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fno-math-errno",
            "-fno-trapping-math", "-mfma", "-fwrapv", "-Wall", "-Wno-unused-function", "-Wno-unused-variable",
            "-Wno-misleading-indentation", "-Wno-parentheses"]

_FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def _strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def _matching_paren(s: str, i: int) -> int:
    """index of the ')' matching the '(' at s[i]"""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses")


def _brace_constructors(s: str) -> str:
    """R11"""
    pat = re.compile(r"(?<![\w.])(vec[234]|Ray)\(")
    pos = 0
    while True:
        m = pat.search(s, pos)
        if not m:
            return s
        open_i = m.end() - 1
        close_i = _matching_paren(s, open_i)
        s = s[:open_i] + "{" + s[open_i + 1:close_i] + "}" + s[close_i + 1:]
        pos = open_i + 1


def _split_top_level(args: str) -> list[str]:
    out, depth, cur = [], 0, []
    for ch in args:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur))
    return out


def _check_evaluation_order(s: str, name: str) -> None:
    """Run on the text BEFORE R11.  Brace lists fix the order inside a constructor; everywhere else an argument list or
    an expression statement holding RNG-advancing calls in two places would have unspecified order in C++ — refuse to build
    rather than guess.  A constructor whose arguments advance the RNG counts as one RNG-advancing call."""
    impure = re.compile(r"\b(" + "|".join(IMPURE) + r"|CTOR_IMPURE)\s*\(")
    ctor = re.compile(r"(?<![\w.])(vec[234]|Ray)\(")
    while True:
        m = ctor.search(s)
        if not m:
            break
        j = _matching_paren(s, m.end() - 1)
        s = s[:m.start()] + ("CTOR_IMPURE()" if impure.search(s[m.end():j]) else "CTOR_PURE") + s[j + 1:]
    for m in re.finditer(r"\(", s):
        try:
            j = _matching_paren(s, m.start())
        except ValueError:
            continue
        parts = _split_top_level(s[m.start() + 1:j])
        if len(parts) > 1 and sum(1 for p in parts if impure.search(p)) > 1:
            raise RuntimeError(f"{name}: two RNG-advancing arguments in one call: {s[m.start():j + 1]!r}")
    for stmt in re.split(r"[;{}]", s):
        if len(impure.findall(stmt)) > 1:
            raise RuntimeError(f"{name}: two RNG-advancing calls in one expression: {stmt.strip()!r}")


def translate(src: str, name: str, capacity=None) -> str:
    s = _strip_comments(src)
    if capacity is not None:                                                                        # R13
        s, n1 = re.subn(r"\bSpheres\[256\]", f"Spheres[{int(capacity[0])}]", s)
        s, n2 = re.subn(r"\bCuboids\[64\]", f"Cuboids[{int(capacity[1])}]", s)
        if name.startswith("PathTracing") and (n1 != 1 or n2 != 1):
            raise RuntimeError(f"{name}: expected exactly one Spheres[256] and one Cuboids[64] declaration")
    s = re.sub(r"^[ \t]*#version[^\n]*", "", s, flags=re.M)                                         # R1
    s = re.sub(r"(\d+\.\d+)\.xxx\b", r"vec3(\1)", s)                                                # R7 (literal swizzle)
    s = re.sub(r"\.(xyz|rgb|xy|zw)\b(?!\s*\()", r".\1()", s)                                        # R7
    s = _FLOAT_LIT.sub(lambda m: "Float(" + m.group(1) + "f)", s)                                   # R2
    s = re.sub(r"\bfloat\b", "Float", s)                                                            # R3
    s = re.sub(r"^[ \t]*layout\s*\([^)]*\)\s*in\s*;", "", s, flags=re.M)                            # R4 (local_size)
    s = re.sub(r"^[ \t]*(?:layout\s*\([^)]*\)\s*)?(uniform|in|out)\s+(\w+)\s*\{",                   # R5 (+R12 for in/out)
               lambda m: ("GLSL_UNIFORM " if m.group(1) == "uniform" else "GLSL_INVOCATION ") + f"struct {m.group(2)} {{", s, flags=re.M)
    s = re.sub(r"^([ \t]*)layout\s*\([^)]*\)\s*", r"\1", s, flags=re.M)                             # R4
    # R4 + R12: remaining global qualifiers.  `uniform` = shared state written by the harness; `in` / `out` = per-invocation.
    s = re.sub(r"\b(?:restrict|writeonly|readonly)\s+", "", s)
    s = re.sub(r"^[ \t]*uniform\s+(\w+)", r"GLSL_UNIFORM \1", s, flags=re.M)
    s = re.sub(r"^[ \t]*(?:in|out)\s+(\w+\s+\w+\s*;)", r"GLSL_INVOCATION \1", s, flags=re.M)
    s = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", s)                                    # R6
    s = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void glsl_main()", s)                                   # R8
    s = re.sub(r"^([ \t]*)Material\s+Material\s*;", r"\1struct Material Material;", s, flags=re.M)  # R9
    s = re.sub(r"\b(\w+)\[(\d+)\]\s+(\w+)\s*;", r"\1 \3[\2];", s)                                   # R10
    _check_evaluation_order(s, name)
    s = _brace_constructors(s)                                                                      # R11
    s = re.sub(r"^(uint|int|Float|vec[234])\s+(\w+)\s*;", r"GLSL_INVOCATION \1 \2;", s, flags=re.M)  # R12 (plain globals)
    return _wrap_invocation(s, name)                                                                # R12


def _top_level_items(s: str) -> list[str]:
    """Split translated text into top-level items: preprocessor lines, `...;` declarations (struct bodies included) and
    function definitions (which end at their closing brace)."""
    items, depth, start, i, n = [], 0, 0, 0, len(s)
    while i < n:
        ch = s[i]
        if depth == 0 and ch == "#" and s[start:i].strip() == "":
            j = s.find("\n", i)
            j = n if j < 0 else j
            items.append(s[i:j])
            start = i = j + 1
            continue
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                head = s[start:i]
                is_function = "(" in head[:head.index("{")] and not re.match(r"\s*(?:GLSL_\w+\s+)?struct\b", head)
                if is_function:
                    items.append(s[start:i + 1])
                    start = i + 1
        elif ch == ";" and depth == 0:
            items.append(s[start:i + 1])
            start = i + 1
        i += 1
    if s[start:].strip():
        raise RuntimeError(f"unterminated top-level item: {s[start:].strip()[:60]!r}")
    return [it.strip() for it in items if it.strip() and it.strip() != ";"]


def _wrap_invocation(s: str, name: str) -> str:
    """R12: GLSL globals are per invocation.  Everything an invocation owns — its mutable globals, its built-in variables and
    all functions (they read those globals) — becomes a member of `struct Invocation`, one object per invocation; types,
    uniforms and uniform blocks stay at namespace scope, shared.  Forward declarations are dropped (members need none).
    GLSL_FN / GLSL_UNIFORM are empty for g++ and `__host__ __device__` / `__constant__` for nvcc (the same text is compiled
    for the GPU by the CUDA harness)."""
    shared, members = [], []
    for it in _top_level_items(s):
        if it.startswith("#"):
            shared.append(it)
        elif it.startswith("GLSL_INVOCATION"):
            members.append("    " + it[len("GLSL_INVOCATION"):].strip())
        elif it.startswith("GLSL_UNIFORM") or re.match(r"struct\b", it):
            shared.append(it)
        elif "{" in it:
            members.append("    GLSL_FN " + it)
        elif re.match(r"[\w&\s]+?\b\w+\s*\([^{}]*\)\s*;$", it, flags=re.S):
            continue                                            # forward declaration
        else:
            raise RuntimeError(f"{name}: cannot classify top-level item {it[:80]!r}")
    return ("\n".join(shared) + "\n\nstruct Invocation {\n    uvec3 gl_GlobalInvocationID;\n    vec4 gl_FragCoord;\n"
            + "\n".join(members) + "\n};\n")


def _sha256(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def shader_paths(reference: str) -> dict[str, str]:
    base = os.path.join(reference, "OpenTK-PathTracer", "res", "shaders")
    return {k: os.path.join(base, v[0]) for k, v in SHADERS.items()}


def capacity_lib(capacity) -> str:
    return os.path.join(OUT_DIR, f"libglsl_ref_{int(capacity[0])}x{int(capacity[1])}.so")


def build(reference: str = "/root/reference", keep: bool = False, verbose: bool = False, alt_model: bool = False, capacity=None) -> str:
    """alt_model=True builds oracle/_ref/libglsl_ref_alt.so instead: the same shaders under a second admissible evaluation
    model (IEEE division, libm transcendentals, nothing fused) — a measuring instrument for tools/model_sensitivity.py,
    never a parity reference."""
    paths = shader_paths(reference)
    lib_out = LIB_ALT if alt_model else (capacity_lib(capacity) if capacity is not None else LIB)
    flags = CXXFLAGS + (["-DGLSL_SHIM_ALT_MODEL"] if alt_model else [])
    for p in paths.values():
        if not os.path.exists(p):
            raise FileNotFoundError(p)
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="glsl_ref_")
    try:
        objs = []
        for key, (rel, harness, ns) in SHADERS.items():
            with open(paths[key], "r", encoding="utf-8-sig") as f:
                text = translate(f.read(), rel, capacity)
            gen = os.path.join(tmp, f"{key}_translated.inc")
            with open(gen, "w") as f:
                f.write(text)
            tu = os.path.join(tmp, f"{key}.cpp")
            with open(tu, "w") as f:
                f.write('#include "glsl_shim.hpp"\n'
                        f"namespace glsl {{ namespace {ns} {{\n"
                        f'#include "{gen}"\n'
                        "}}\n"
                        f'#include "{harness}"\n')
            obj = os.path.join(tmp, f"{key}.o")
            cmd = [CXX, *flags, "-I", HERE, "-c", tu, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
            objs.append(obj)
        subprocess.run([CXX, "-shared", "-fopenmp", "-o", lib_out, *objs, "-lm"], check=True)
        if alt_model or capacity is not None:
            return lib_out
        manifest = {
            "built_from": {k: {"path": os.path.relpath(p, reference), "sha256": _sha256(p)} for k, p in paths.items()},
            "shim": {n: _sha256(os.path.join(HERE, n)) for n in
                     ("glsl_shim.hpp", "glsl_model.h", "build_ref.py", *(v[1] for v in SHADERS.values()))},
            "cxx": subprocess.run([CXX, "--version"], capture_output=True, text=True).stdout.splitlines()[0],
            "flags": CXXFLAGS,
        }
        with open(MANIFEST, "w") as f:
            json.dump(manifest, f, indent=1)
    finally:
        if keep:
            print(f"translated text kept in {tmp}", file=sys.stderr)
        else:
            shutil.rmtree(tmp, ignore_errors=True)
    return lib_out


def cuda_lib(fast: bool = False, capacity=None) -> str:
    tag = ("_fast" if fast else "") + (f"_{int(capacity[0])}x{int(capacity[1])}" if capacity is not None else "")
    return os.path.join(OUT_DIR, f"libglsl_ref_cuda{tag}.so")


def build_cuda(reference: str = "/root/reference", fast: bool = False, capacity=None, verbose: bool = False) -> str:
    """The same translated compute.glsl, compiled by nvcc for sm_100a and dispatched in the reference's own launch shape
    (oracle/ref_harness_pt_cuda.inc): the GL-compute proxy built from the reference's source.  fast=False: the evaluation
    model of glsl_model.h (-fmad=false), must equal the oracle bit for bit.  fast=True: -use_fast_math + MUFU built-ins +
    contraction, roughly a GL driver's code generation; timing only."""
    path = shader_paths(reference)["pt"]
    os.makedirs(OUT_DIR, exist_ok=True)
    out = cuda_lib(fast, capacity)
    tmp = tempfile.mkdtemp(prefix="glsl_ref_cuda_")
    try:
        with open(path, "r", encoding="utf-8-sig") as f:
            text = translate(f.read(), SHADERS["pt"][0], capacity)
        gen = os.path.join(tmp, "pt_translated.inc")
        with open(gen, "w") as f:
            f.write(text)
        tu = os.path.join(tmp, "pt.cu")
        with open(tu, "w") as f:
            f.write('#include <cstdio>\n#include "glsl_shim.hpp"\nnamespace glsl { namespace pt {\n'
                    f'#include "{gen}"\n' "}}\n" '#include "ref_harness_pt_cuda.inc"\n')
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
               "-cudart", "static", "-I", HERE]
        cmd += ["-use_fast_math", "-DGLSL_SHIM_FAST_GPU"] if fast else ["-fmad=false"]
        if capacity is not None:
            cmd += ["-DGLSL_UNIFORM_DEVICE"]
        cmd += ["-o", out, tu]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
            print(" ".join(cmd))
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--keep", action="store_true", help="leave the translated text in its /tmp directory")
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("--alt-model", action="store_true", help="build libglsl_ref_alt.so (second evaluation model, for tools/model_sensitivity.py)")
    ap.add_argument("--capacity", type=int, nargs=2, metavar=("SPHERES", "CUBOIDS"), help="R13: build libglsl_ref_<S>x<C>.so with larger UBO arrays (BASELINE config 3)")
    ap.add_argument("--cuda", action="store_true", help="compile compute.glsl with nvcc instead (libglsl_ref_cuda[_fast].so, the GL-compute proxy)")
    ap.add_argument("--fast", action="store_true", help="with --cuda: -use_fast_math + MUFU built-ins (timing only)")
    a = ap.parse_args()
    if a.cuda:
        print(build_cuda(a.reference, a.fast, a.capacity, a.verbose))
    else:
        print(build(a.reference, a.keep, a.verbose, a.alt_model, a.capacity))
