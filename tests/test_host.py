"""CPU tests of the host side: scene byte producers, the C-ABI surface, error behaviour, the tile partition and its
gather (gloo, world_size 2).  No compute calls into the CUDA library are made here."""
import ctypes as C
import json
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, f32_same

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PINS = json.load(open(os.path.join(GOLD, "layout_pins.json")))


# ------------------------------------------------------------------------------- std140 layout
def test_layout_pins(ptb):
    sc = ptb.scene
    assert sc.MATERIAL_SIZE == PINS["material"] and sc.SPHERE_SIZE == PINS["sphere"] and sc.CUBOID_SIZE == PINS["cuboid"]
    assert sc.BASIC_DATA_SIZE == PINS["basic_data_ubo"]
    scene = sc.load_default_scene()
    assert scene.ubo_size == PINS["game_objects_ubo"]
    assert scene.cuboids[0].BufferOffset == PINS["cuboid_base"]
    assert scene.spheres[3].BufferOffset == 3 * 80 and scene.cuboids[2].BufferOffset == 20480 + 2 * 96
    assert len(sc.atmosphere_ubo_bytes()) == PINS["atmosphere_ubo"]


def test_material_packing_and_clamps(ptb):
    sc = ptb.scene
    m = sc.Material.new((0.1, 0.2, 0.3), (1, 2, 3), (4, 5, 6), specularChance=1.5, specularRoughness=0.25, indexOfRefraction=0.5,
                        refractionChance=0.9, refractionRoughnes=0.75)
    assert m.SpecularChance == 1.0 and m.IOR == 1.0 and m.RefractionChance == 0.0      # Material.cs:26-29
    d = m.GetGPUFriendlyData()
    assert d.shape == (4, 4) and d.dtype == np.float32
    assert d[0].tolist() == [np.float32(0.1), np.float32(0.2), np.float32(0.3), 1.0]    # Albedo, SpecularChance
    assert d[1].tolist() == [1.0, 2.0, 3.0, 0.25]                                          # Emissiv, SpecularRoughness
    assert d[2].tolist() == [4.0, 5.0, 6.0, 0.0]                                           # Absorbance, RefractionChance
    assert d[3].tolist() == [0.75, 1.0, 0.0, 0.0]                                          # RefractionRoughness, IOR
    z = sc.Material.Zero()
    z.SpecularChance = np.float32(7.0)    # field writes bypass the ctor clamps, as in MainWindow.cs:225-229
    assert z.GetGPUFriendlyData()[0, 3] == 7.0


def test_default_scene_pins(ptb):
    scene = ptb.load_default_scene()
    assert len(scene.spheres) == PINS["default_spheres"] and len(scene.cuboids) == PINS["default_cuboids"]
    grid = scene.spheres[:36]
    assert all(s.Position[2] == -5 and s.Radius == np.float32(1.3) for s in grid)
    assert grid[0].Material.SpecularChance == 0.0 and grid[35].Material.SpecularChance == 1.0
    assert grid[7].Material.SpecularChance == np.float32(1) / np.float32(5) and grid[7].Material.SpecularRoughness == np.float32(1) / np.float32(5)
    glass = scene.spheres[36:]
    assert all(s.Position[2] == -20 for s in glass) and {float(s.Position[1]) for s in glass} == {3.0, -6.0}
    assert all(s.Material.RefractionChance == np.float32(0.98) and s.Material.SpecularChance == np.float32(0.02) for s in glass)
    light = scene.cuboids[1]
    assert np.allclose(light.Material.Emissiv, np.array([0.917, 0.945, 0.513], np.float32) * np.float32(5))
    assert abs(float(light.Position[1]) - 18.49) < 1e-5
    assert scene.cuboids[3].Material.IOR == 1.0 and scene.cuboids[3].Material.RefractionChance == np.float32(0.954)    # glass front wall
    assert scene.cuboids[4].Material.SpecularChance == 1.0                                                                # mirror-ish right wall
    # Min/Max, not position/dimensions, are uploaded (Cuboid.cs:23-30)
    d = scene.cuboids[6].GetGPUFriendlyData()
    assert np.allclose(d[0, :3], [-16.5, -13.495, -16.5]) and np.allclose(d[1, :3], [-13.5, -7.495, -13.5])


def test_scene_bytes_match_golden(ptb):
    ubo = np.frombuffer(ptb.load_default_scene().ubo_bytes(), dtype=np.uint8)
    assert (ubo[:48 * 80] == np.load(os.path.join(GOLD, "default_scene_ubo.npy"))).all()
    assert (ubo[20480:20480 + 7 * 96] == np.load(os.path.join(GOLD, "default_scene_cuboids.npy"))).all()
    assert (ubo[48 * 80:20480] == 0).all()
    basic = np.frombuffer(ptb.scene.basic_data_bytes(ptb.default_camera(), 64, 64), dtype=np.uint8)
    assert (basic == np.load(os.path.join(GOLD, "default_basic_ubo_64x64.npy"))).all()


def test_synthetic_scene_is_deterministic_and_sized(ptb):
    a = ptb.synthetic_scene(64, 16, seed=1234)
    b = ptb.synthetic_scene(64, 16, seed=1234)
    assert a.ubo_bytes() == b.ubo_bytes()
    assert len(a.spheres) == 64 and len(a.cuboids) == 16 and a.ubo_size == 64 * 80 + 16 * 96
    assert a.cuboids[0].BufferOffset == 64 * 80      # cuboid base follows the sphere CAPACITY (Cuboid.cs:21)
    assert all(s.Material.RefractionRoughnes >= np.float32(0.05) for s in a.spheres)
    assert all(0 <= s.Material.RefractionChance <= 1 - s.Material.SpecularChance for s in a.spheres)


def test_host_buffer_subdata_semantics(ptb):
    buf = ptb.scene.HostBuffer(32)
    buf.SubData(8, 8, np.array([1.0, 2.0], np.float32))
    assert np.frombuffer(buf.bytes(), np.float32)[2:4].tolist() == [1.0, 2.0]
    buf.SubData(0, 16, np.array([7.0, 8.0, 9.0], np.float32))      # the ViewPos write passes 12 bytes with size 16 (MainWindow.cs:132)
    assert np.frombuffer(buf.bytes(), np.float32)[:4].tolist() == [7.0, 8.0, 9.0, 0.0]
    with pytest.raises(ValueError):
        buf.SubData(24, 16, b"\0" * 16)


# ------------------------------------------------------------------------------- OpenTK math restatement
def test_opentk_matrices(ptb):
    sc = ptb.scene
    cam = ptb.default_camera()
    v = cam.View
    assert np.allclose(v[:3, :3] @ v[:3, :3].T, np.eye(3), atol=1e-6)           # orthonormal basis
    assert np.allclose(sc.inverted(v) @ v, np.eye(4), atol=1e-5)
    p = sc.create_perspective_fov(sc.degrees_to_radians(90.0), 1.0, 0.1, 10.0)
    assert np.allclose([p[0, 0], p[1, 1]], [1.0, 1.0], atol=1e-6) and p[2, 3] == -1 and p[3, 3] == 0
    assert np.allclose(sc.inverted(p) @ p, np.eye(4), atol=1e-5)
    # row-vector convention: the eye maps to the origin, the view direction to -z
    eye = np.append(cam.Position, 1).astype(np.float32)
    assert np.allclose(eye @ v, [0, 0, 0, 1], atol=1e-5)
    ahead = np.append(cam.Position + cam.ViewDir, 1).astype(np.float32)
    assert np.allclose(ahead @ v, [0, 0, -1, 1], atol=1e-5)
    with pytest.raises(ValueError):
        sc.inverted(np.zeros((4, 4), np.float32))


def test_atmosphere_inputs(ptb):
    lp = ptb.scene.atmosphere_light_pos(0.5)
    assert lp[0] == 0 and abs(lp[2] + 1.496e11) < 1e5 and abs(lp[1]) < 2e4     # sun on the -z horizon
    ubo = np.frombuffer(ptb.scene.atmosphere_ubo_bytes(), np.float32)
    inv_view_px = ubo[16:32].reshape(4, 4)
    d = np.array([0, 0, -1, 0], np.float32) @ inv_view_px                       # the +X face looks down +x
    assert np.allclose(d[:3], [1, 0, 0], atol=1e-6)


# ------------------------------------------------------------------------------- C ABI surface
def _header_symbols():
    text = open(os.path.join(ROOT, "include", "ptb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ptb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(ptb):
    from importlib import import_module
    _lib = import_module("opentk-pathtracer_b200._lib")
    syms = _header_symbols()
    assert len(syms) >= 35
    assert sorted(_lib.SIGNATURES) == syms           # the ctypes binding declares exactly the header's entry points
    L = _lib.load()
    for s in syms:
        assert hasattr(L, s), f"libptb200.so does not export {s}"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (ptb_[a-z0-9_]+)", out))
    assert exported == set(syms)                      # and nothing undeclared leaks out
    assert L.ptb_version() == 100


def test_csharp_shim_binds_the_declared_abi():
    """csharp/PtbNative.cs (the P/Invoke half of the drop-in for src/Render/PathTracer.cs) cannot be compiled here (no .NET):
    every [DllImport] must at least name an entry point of include/ptb200.h with the same number of arguments, and the shim
    class must keep the reference class's public members (PathTracer.cs:11-140)."""
    hdr = open(os.path.join(ROOT, "include", "ptb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    c_args = {}
    for m in re.finditer(r"\b(?:int|void|float|const char\*)\s+(ptb_[a-z0-9_]+)\(([^)]*)\);", hdr):
        args = m.group(2).strip()
        c_args[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    cs = open(os.path.join(ROOT, "csharp", "PtbNative.cs")).read()
    imports = re.findall(r"\[DllImport\(Lib\)\]\s*public static extern (?:unsafe )?\w+ (ptb_[a-z0-9_]+)\(([^)]*)\);", cs)
    assert len(imports) >= 25
    for name, args in imports:
        assert name in c_args, f"{name} is not declared in ptb200.h"
        n = 0 if not args.strip() else len(args.split(","))
        assert n == c_args[name], f"{name}: {n} arguments in C#, {c_args[name]} in C"
    shim = open(os.path.join(ROOT, "csharp", "PathTracer.cs")).read()
    for member in ("NumSpheres", "NumCuboids", "RayDepth", "SPP", "FocalLength", "ApertureDiameter", "EnvironmentMap", "Result", "Samples",
                   "void Render()", "void SetSize(int width, int height)", "void ResetRenderer()"):
        assert member in shim, member
    used = set(re.findall(r"Ptb\.(ptb_[a-z0-9_]+)", shim))
    assert used <= {n for n, _ in imports}, used - {n for n, _ in imports}
    assert {"ptb_register_gl_texture", "ptb_present_gl", "ptb_render"} <= used      # Result goes to GL on the device, no host round trip
    assert "ptb_read_result" not in used


def test_library_has_sm100a_code_only(ptb):
    from importlib import import_module
    _lib = import_module("opentk-pathtracer_b200._lib")
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "opentk-pathtracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pt_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
                assert "glsl_ref" not in text and "glsl_shim" not in text and "build_ref" not in text and "glref_" not in text, f
    # and the library the product loads links nothing of the checkers
    import subprocess
    from importlib import import_module
    lib = import_module("opentk-pathtracer_b200._lib").lib_path()
    needed = subprocess.run(["readelf", "-d", lib], capture_output=True, text=True).stdout
    assert "oracle" not in needed and "glsl_ref" not in needed


def test_fails_loudly_without_a_gpu(ptb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ptb.PtbError) as e:
        ptb.PathTracer(None, 64, 64, 13, 1, 20.0, 0.14)
    assert "error -2" in str(e.value)          # PTB_E_CUDA, no silent CPU fallback
    L = ptb.load_library()
    assert L.ptb_set_spp(None, 1) == -1 and b"null" in L.ptb_last_error()


# ------------------------------------------------------------------------------- tile partition + gather
def test_partition_covers_every_row_once(ptb):
    from importlib import import_module
    D = import_module("opentk-pathtracer_b200.distributed")
    for height, world, stripe in [(1080, 8, 8), (1080, 3, 8), (2160, 8, 16), (123, 4, 8), (7, 8, 8), (64, 1, 8)]:
        rows = np.concatenate([D.local_rows_of(r, world, stripe, height) for r in range(world)])
        assert sorted(rows.tolist()) == list(range(height))
        assert D.max_local_rows(world, stripe, height) == max(D.local_row_count(r, world, stripe, height) for r in range(world))
        g = np.zeros((world, D.max_local_rows(world, stripe, height), 5, 1), np.float32)
        for r in range(world):
            lr = D.local_rows_of(r, world, stripe, height)
            g[r, :lr.size, :, 0] = lr[:, None]
        full = D.deinterleave_host(g, height, world, stripe)
        assert (full[:, 0, 0] == np.arange(height)).all()


_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
import ptb200
from importlib import import_module
D = import_module("opentk-pathtracer_b200.distributed")
from oracle import oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
sc = ptb200.scene
W, H, stripe = 40, 52, 8
env = O.atmosphere(8, sc.atmosphere_ubo_bytes(), sc.atmosphere_light_pos(0.5), 15.0, 4, 2)
scene, cam = sc.load_default_scene(), sc.default_camera()
basic, ubo = sc.basic_data_bytes(cam, W, H), scene.ubo_bytes()
rows = D.local_rows_of(rank, world, stripe, H)
local = np.zeros((D.max_local_rows(world, stripe, H), W, 4), np.float32)
scratch = np.zeros((H, W, 4), np.float32)
for frame in range(3):
    # this rank renders ONLY its stripes (global pixel coordinates seed the RNG), keeping its running mean locally
    scratch[rows] = local[:rows.size]
    for s in range(0, rows.size, stripe):
        O.render(scratch, basic, ubo, env, frame=frame, spp=1, ray_depth=13, focal_length=20.0, aperture_diameter=0.14,
                 n_spheres=48, n_cuboids=7, rows=(int(rows[s]), int(rows[min(s + stripe, rows.size) - 1]) + 1), n_threads=1)
    local[:rows.size] = scratch[rows]
    gathered, work = D.gather_stripes(torch.from_numpy(local), world, dst=0)
    if rank == 0:
        full = D.deinterleave_host(gathered.numpy(), H, world, stripe)
if rank == 0:
    ref = np.zeros((H, W, 4), np.float32)
    for frame in range(3):
        O.render(ref, basic, ubo, env, frame=frame, spp=1, ray_depth=13, focal_length=20.0, aperture_diameter=0.14, n_spheres=48, n_cuboids=7)
    assert (full.view(np.uint32) == ref.view(np.uint32)).all(), "gathered tiles differ from the single-process render"
    print("GATHER_OK")
dist.barrier()
dist.destroy_process_group()
"""


_SHARED_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from importlib import import_module
D = import_module("opentk-pathtracer_b200.distributed")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
W, H, stripe = 24, 53, 8                       # 7 stripes, the last one ragged (5 rows)
shared = D.SharedHostFrame(H * W * 12, 2, rank, world, register=False)
assert shared.path and not os.path.exists(shared.path)        # unlinked once everybody mapped it
truth = np.arange(H * W * 3, dtype=np.float32).reshape(H, W, 3)
rows = D.local_rows_of(rank, world, stripe, H)
for k in (0, 1, 2):                             # buffer index wraps modulo 2
    local = np.zeros((D.max_local_rows(world, stripe, H), W, 3), np.float32)
    local[:rows.size] = truth[rows] + k
    frame = shared.view(k, (H, W, 3), torch.float32).numpy()
    D.scatter_rows_host(local, frame, rank, world, stripe)
    dist.barrier()
    assert (frame == truth + k).all(), "rows written by the other rank are not visible through the shared mapping"
    dist.barrier()
assert shared.ptr(2) == shared.ptr(0) and shared.ptr(1) - shared.ptr(0) >= H * W * 12
shared.close()
if rank == 0:
    print("SHARED_OK")
dist.destroy_process_group()
"""


def test_two_rank_shared_host_frame(tmp_path):
    """The host side of the N>1 read-back on CPU: two gloo processes set up one shared frame mapping (no CUDA pinning here),
    each writes ITS rows with the scatter plan the C ABI uses, and both see the complete frame."""
    script = tmp_path / "shared_worker.py"
    script.write_text(_SHARED_WORKER.format(root=ROOT))
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "SHARED_OK" in outs[0]


def test_scatter_plan_covers_every_row_once(ptb):
    from importlib import import_module
    D = import_module("opentk-pathtracer_b200.distributed")
    for height, world, stripe in [(1080, 8, 8), (1080, 3, 8), (2160, 8, 16), (203, 8, 16), (203, 4, 5), (7, 8, 8), (64, 1, 8)]:
        frame = np.full((height, 3), -1.0, np.float32)
        for r in range(world):
            rows = D.local_rows_of(r, world, stripe, height)
            local = np.zeros((D.max_local_rows(world, stripe, height), 3), np.float32)
            local[:rows.size, 0] = rows
            D.scatter_rows_host(local, frame, r, world, stripe)
        assert (frame[:, 0] == np.arange(height)).all(), (height, world, stripe)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo_gather_equals_single_render(tmp_path):
    """The N>1 host path on CPU: two processes, gloo, each rendering its own stripes (the oracle stands in for the GPU
    kernel), one gather per frame, de-interleave on rank 0 == unpartitioned render, bitwise."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "GATHER_OK" in outs[0]


def test_cpu_picking_matches_the_shader_fold(ptb, oracle):
    """MainWindow.RayTrace (the mouse-picking fold on the host) mirrored in scene.pick.  It is the same
    order-dependent fold as compute.glsl's RayTrace, in C# float arithmetic (IEEE division instead of a*rcp(b)), so it must
    select the same primitive as the oracle for every cursor position, with T equal up to the rounding of the discriminant."""
    sc = ptb.scene
    scene, cam = sc.load_default_scene(), sc.default_camera()
    W, H = 1280, 720
    rng = np.random.default_rng(9)
    rays, picks = [], []
    for _ in range(400):
        x, y = int(rng.integers(0, W)), int(rng.integers(0, H))
        obj, t1, t2, ray = sc.pick_at_cursor(scene, cam, W, H, x, y)
        rays.append(np.concatenate([ray.Origin, ray.Direction]))
        picks.append((obj, t1, t2))
    ref = oracle.ray_trace(np.asarray(rays, np.float32), scene.ubo_bytes(), 256, 48, 7)
    objs = scene.objects()
    kinds = set()
    for (obj, t1, t2), r in zip(picks, ref):
        assert (obj is not None) == bool(r[0])
        if obj is None:
            continue
        kinds.add(type(obj).__name__)
        T = t2 if t1 < 0 else t1
        assert abs(float(T) - float(r[1])) <= 2e-4 * max(1.0, abs(float(r[1])))     # unfused C# dot products vs the shader model's fma chains, amplified by b*b - c
        assert np.float32(obj.Material.Albedo[0]) == r[10] and np.float32(obj.Material.Emissiv[0]) == r[11]
    assert kinds == {"Sphere", "Cuboid"}
    # known answers: a ray down the axis of sphere 0 hits it at |o - c| -/+ r; from inside a cuboid t1 < 0 < t2
    s0 = scene.spheres[0]
    o = (np.asarray(s0.Position, np.float32) + np.array([0, 0, 10], np.float32)).astype(np.float32)
    obj, t1, t2 = sc.pick(sc.Scene(spheres=[s0]), sc.Ray(o, np.array([0, 0, -1], np.float32)))
    assert obj is s0 and abs(float(t1) - (10 - float(s0.Radius))) < 1e-5 and abs(float(t2) - (10 + float(s0.Radius))) < 1e-5
    room = scene.cuboids[0]
    centre = ((room.Min + room.Max) * np.float32(0.5)).astype(np.float32)
    hit, t1, t2 = sc.cuboid_intersects_ray(room, sc.Ray(centre, sc._normalize(np.array([0.3, 1.0, 0.2], np.float32))))
    assert hit and t1 < 0 < t2
    assert sc.pick(sc.Scene(), sc.Ray(o, np.array([0, 0, -1], np.float32)))[0] is None


def test_skybox_image_loading(ptb, tmp_path):
    """Helper.ParallelLoadCubemapImages mirrored in scene.load_cubemap_images: face order, row order, RGBA expansion of every
    PNG colour type, and the reference's error cases."""
    from PIL import Image
    sc = ptb.scene
    rng = np.random.default_rng(3)
    faces = rng.integers(0, 256, (6, 8, 8, 4), dtype=np.uint8)
    modes = ["RGBA", "RGB", "RGBA", "L", "RGB", "P"]
    paths, want = [], []
    for i, (name, mode) in enumerate(zip(sc.SKYBOX_FACE_FILES, modes)):
        im = Image.fromarray(faces[i], "RGBA").convert(mode)
        p = tmp_path / (name.upper() if i % 2 else name)            # mixed case on disk, like posX.png in the repository
        im.save(p, format="PNG")
        paths.append(str(p))
        want.append(np.asarray(im.convert("RGBA")))
    got = sc.load_cubemap_images(sc.skybox_paths(str(tmp_path)))
    assert got.shape == (6, 8, 8, 4) and got.dtype == np.uint8 and got.flags.c_contiguous
    assert (got == np.stack(want)).all()
    assert (got[0] == faces[0]).all() and (got[1][..., :3] == faces[1][..., :3]).all() and (got[1][..., 3] == 255).all()
    with pytest.raises(ValueError, match="six"):
        sc.load_cubemap_images(paths[:5])
    with pytest.raises(FileNotFoundError):
        sc.load_cubemap_images(paths[:5] + [str(tmp_path / "missing.png")])
    Image.fromarray(faces[0][:, :4], "RGBA").save(tmp_path / "wide.png")
    with pytest.raises(ValueError, match="squares"):
        sc.load_cubemap_images(paths[:5] + [str(tmp_path / "wide.png")])


@pytest.mark.skipif(not os.path.isdir("/root/reference/OpenTK-PathTracer/res/textures/EnvironmentMap"), reason="the reference's skybox images are not mounted")
def test_the_references_own_skybox_through_oracle_and_compiled_shader(ptb, oracle):
    """The six 2048^2 PNG faces the reference ships, loaded like Helper.cs does, decoded like an Srgb8Alpha8 texture, used as
    the EnvironmentMap: the oracle and the compiled reference shader render the same bits with it."""
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref not built")
    sc = ptb.scene
    faces = sc.load_cubemap_images(sc.skybox_paths("/root/reference/OpenTK-PathTracer/res/textures/EnvironmentMap"))
    assert faces.shape == (6, 2048, 2048, 4) and (faces[..., 3] == 255).all()
    small = np.ascontiguousarray(faces[:, ::8, ::8])                 # 256^2 faces keep the test light; decoding is per texel
    env = oracle.srgb8_to_linear(small)
    assert env.shape == (6, 256, 256, 4) and 0.0 < float(env[..., :3].mean()) < 1.0
    scene, cam = sc.load_default_scene(), sc.default_camera()
    W, H = 96, 54
    basic, ubo = sc.basic_data_bytes(cam, W, H), scene.ubo_bytes()
    a, b = np.zeros((H, W, 4), np.float32), np.zeros((H, W, 4), np.float32)
    for f in range(2):
        kw = dict(frame=f, spp=2, ray_depth=13, focal_length=20.0, aperture_diameter=0.14, n_spheres=48, n_cuboids=7)
        oracle.render(a, basic, ubo, env, **kw)
        R.render(b, basic, ubo, env, **kw)
    assert (a.view(np.uint32) == b.view(np.uint32)).all()


def test_atmospheric_scatterer_mirror_defaults_and_clamp(ptb):
    """AtmosphericScatterer.cs:11-57,91-94 — property surface without a GPU (Render() is exercised by the GPU tests)."""
    class FakeTracer:
        def GenerateAtmosphere(self, *a, fast=False):
            self.args = a
    t = FakeTracer()
    a = ptb.AtmosphericScatterer(t, 256)
    assert (a.Size, a.Time, a.ISteps, a.JSteps, a.LightIntensity) == (256, 0.5, 50, 15, 15.0)
    a.LightIntensity = -3.0
    assert a.LightIntensity == 0.0                      # Math.Max(value, 0.0f)
    a.Time, a.ISteps, a.JSteps, a.LightIntensity = 0.25, 20, 5, 22.0
    a.SetSize(64)
    a.Render()
    assert t.args == (64, 20, 5, 0.25, 22.0)
    lp = a.LightPos
    assert lp.dtype == np.float32 and lp[0] == 0 and abs(float(lp[1]) - 149600000e3) < 1e6 and abs(float(lp[2])) < 1e5   # noon: sun at the zenith
    with pytest.raises(ValueError):
        ptb.AtmosphericScatterer(t, 0)


# ------------------------------------------------------------------------------- fused exchange: the host-side slot bookkeeping
class _FakeExchangeLib:
    """Stands in for libptb200's exchange entry points (include/ptb200.h: ptb_exchange_pending / _acquire / _release /
    _root, ptb_render_frames' refusal when a call would wait for its own releases) so TiledPathTracer's chunking can be
    checked without a GPU.  Frame q is assembled on rank q % world with rotating roots, on rank 0 otherwise."""

    def __init__(self, rank, world, slots, rotate):
        self.rank, self.world, self.slots, self.rotate = rank, world, slots, rotate
        self.seq = 0               # frames rendered so far (every rank renders every frame)
        self.owned = []            # frames this rank is the root of, rendered and not yet acquired
        self.held = None           # the acquired frame
        self.in_ring = 0           # own frames occupying a slot (rendered, not yet released)
        self.log = []

    def root_of(self, q):
        return q % self.world if self.rotate else 0

    def render(self, n):
        mine = [q for q in range(self.seq, self.seq + n) if self.root_of(q) == self.rank]
        if self.in_ring + len(mine) > self.slots:
            return -3              # PTB_E_STATE: the call would wait for releases that cannot be enqueued before it returns
        self.log.append(("render", n))
        self.seq += n
        self.owned += mine
        self.in_ring += len(mine)
        return 0

    def ptb_exchange_pending(self, ctx):
        return len(self.owned)

    def ptb_exchange_acquire(self, ctx, out):
        assert self.held is None and self.owned
        self.held = self.owned.pop(0)
        out._obj.value = 0x1000 + (self.held // (self.world if self.rotate else 1)) % self.slots
        return 0

    def ptb_exchange_release(self, ctx):
        assert self.held is not None
        self.log.append(("release", self.held))
        self.held = None
        self.in_ring -= 1
        return 0

    def ptb_exchange_root(self, ctx, q):
        return self.root_of(self.seq - 1 if q < 0 else q)


@pytest.mark.parametrize("rank,world,slots,rotate,frames", [(1, 2, 4, True, 37), (0, 8, 8, True, 200), (0, 4, 32, False, 100),
                                                            (3, 4, 32, False, 100), (2, 3, 1, True, 7)])
def test_step_batch_never_outruns_the_slot_ring(ptb, rank, world, slots, rotate, frames):
    """distributed.TiledPathTracer.step_batch: chunks are sized so that no Render() call waits for a release that is only
    enqueued after it returns (ptb_abi.cu ptb_render_frames), every frame this rank is the root of is
    handed to the consumer exactly once and in order, and last_root() names the rank holding the newest frame."""
    import importlib
    D = importlib.import_module(ptb.__name__ + ".distributed")
    lib = _FakeExchangeLib(rank, world, slots, rotate)

    class Tracer:
        _L, _ctx = lib, None

        def Render(self, n=1):
            rc = lib.render(n)
            assert rc == 0, "TiledPathTracer asked for more frames than the slot ring can hold"

    tp = object.__new__(D.TiledPathTracer)
    tp.tracer, tp.rank, tp.world, tp.fused, tp.rotate, tp.slots = Tracer(), rank, world, True, rotate, slots
    tp.height, tp.width, tp.channels, tp.device = 2, 2, 3, None
    tp._last_full = None
    tp._slot_tensors = {0x1000 + s: ("slot", s) for s in range(slots)}      # pre-wrapped: no CUDA array interface on CPU
    seen = []
    last = tp.step_batch(frames, consumer=seen.append)
    owned = [q for q in range(frames) if lib.root_of(q) == rank]
    assert [e[1] for e in lib.log if e[0] == "release"] == owned
    assert len(seen) == len(owned) and lib.in_ring == 0 and lib.seq == frames
    if owned:
        assert last == seen[-1] == ("slot", (owned[-1] // (world if rotate else 1)) % slots)
    else:
        assert last is None
    assert max(n for kind, n in lib.log if kind == "render") <= slots * (world if rotate else 1)
    assert tp.last_root() == lib.root_of(frames - 1)
    tp.render(5)                                                            # render() goes through the same chunking when fused
    assert lib.seq == frames + 5 and lib.in_ring == 0
