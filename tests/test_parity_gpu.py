"""GPU parity tests (run on a B200 with `-m gpu`): the CUDA path, called through the C ABI (ctypes -> libptb200.so),
against the CPU oracle on the same seeded inputs.  The bar is BIT-EXACT float32 (any NaN == any NaN): a Monte-Carlo
path is chaotic, so anything looser would hide real divergence.  Tolerance stated once: 0 ulp.
Nothing here reads /root/reference."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import f32_same, oracle_render

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_tracer(ptb, env, W, H, scene, camera, *, spp=1, depth=13, focal=20.0, aperture=0.14, kernel=0):
    pt = ptb.PathTracer(env, W, H, depth, spp, focal, aperture, max_spheres=scene.max_spheres, max_cuboids=scene.max_cuboids)
    pt.LoadScene(scene)
    pt.SetCamera(camera)
    pt.SetKernel(kernel)
    return pt


def assert_same(got, ref, what=""):
    bad = ~f32_same(got, ref)
    assert not bad.any(), f"{what}: {int(bad.any(axis=-1).sum())} differing pixels; first at {np.argwhere(bad)[0].tolist()}"


@pytest.fixture(scope="module")
def tracer256(ptb, env256, default_scene, camera):
    pt = make_tracer(ptb, env256, 256, 256, default_scene, camera)
    yield pt
    pt.Dispose()


# ------------------------------------------------------------------------------- the library actually loaded
def test_native_library_is_loaded(ptb):
    L = ptb.load_library()
    assert L.ptb_version() == 100
    maps = open("/proc/self/maps").read()
    assert "libptb200.so" in maps


# ------------------------------------------------------------------------------- unit probes (ptb_debug_eval)
def test_pcg_stream_matches_golden(tracer256, oracle):
    import json
    for k in json.load(open(os.path.join(GOLD, "pcg_kats.json"))):
        seed = np.array([k["seed"]], np.uint32).view(np.float32)
        got = tracer256.DebugEval(2, seed, 8, 8)
        assert [int(v) for v in got.view(np.uint32)] == k["floats_hex"]


def test_transcendentals_bit_exact(tracer256, oracle):
    x = np.concatenate([np.linspace(0, 6.2831855, 1000003, dtype=np.float32), np.linspace(-50, 50, 100001, dtype=np.float32)])
    s, c = oracle.sincos(x)
    d = tracer256.DebugEval(0, x, x.size, 2 * x.size).reshape(-1, 2)
    assert_same(d[:, 0], s, "sin")
    assert_same(d[:, 1], c, "cos")
    x = np.concatenate([np.linspace(-110, 90, 1000003, dtype=np.float32), np.array([np.nan, np.inf, -np.inf, 0.0, -0.0], np.float32)])
    assert_same(tracer256.DebugEval(1, x, x.size, x.size), oracle.exp(x), "exp")


def test_min_max_rcp_sqrt_model(tracer256):
    rng = np.random.default_rng(1)
    ab = (rng.standard_normal((200000, 2)) * 10.0 ** rng.integers(-30, 30, (200000, 2))).astype(np.float32)
    ab[:8] = np.array([[0.0, -0.0], [-0.0, 0.0], [np.nan, 1], [1, np.nan], [np.inf, 1], [0.0, 4.0], [-0.0, 0.0], [np.nan, np.nan]], np.float32)
    r = tracer256.DebugEval(5, ab, ab.shape[0], 4 * ab.shape[0]).reshape(-1, 4)
    # min/max: NaN loses, -0 < +0 (the oracle's g_min/g_max)
    assert r[0, 0].view(np.uint32) == 0x80000000 and r[1, 0].view(np.uint32) == 0x80000000
    assert r[0, 1].view(np.uint32) == 0 and r[1, 1].view(np.uint32) == 0
    assert r[2, 0] == 1 and r[3, 0] == 1 and r[2, 1] == 1 and r[3, 1] == 1 and np.isnan(r[7, 0])
    with np.errstate(all="ignore"):
        assert_same(r[:, 2], np.float32(1) / ab[:, 0], "rcp")       # correctly rounded reciprocal
        assert_same(r[:, 3], np.sqrt(ab[:, 1]), "sqrt")


def test_cubemap_lookup_bit_exact(tracer256, oracle, env256):
    rng = np.random.default_rng(2)
    dirs = rng.standard_normal((300000, 3)).astype(np.float32)
    dirs[:6] = np.eye(3, dtype=np.float32).repeat(2, 0) * np.array([1, -1] * 3, np.float32)[:, None]
    dirs[6:14] = np.array([[sx, sy, sz] for sx in (1, -1) for sy in (1, -1) for sz in (1, -1)], np.float32)   # corners
    e = rng.standard_normal((3000, 3)).astype(np.float32)
    e[:, 0] = np.sign(e[:, 0]); e[:, 1] = np.sign(e[:, 1])                                                     # exact edges
    dirs[14:3014] = e
    dirs[3014] = np.nan; dirs[3015] = 0; dirs[3016] = (np.inf, 1, 1)
    got = tracer256.DebugEval(3, dirs, dirs.shape[0], 3 * dirs.shape[0]).reshape(-1, 3)
    assert_same(got, oracle.texture_cube(env256, dirs), "texture(samplerCube)")


def test_small_cubemaps_bit_exact(ptb, oracle, default_scene, camera):
    rng = np.random.default_rng(4)
    dirs = rng.standard_normal((20000, 3)).astype(np.float32)
    for n in (1, 2, 3, 16):
        env = rng.random((6, n, n, 4)).astype(np.float32)
        pt = make_tracer(ptb, env, 16, 16, default_scene, camera)
        got = pt.DebugEval(3, dirs, dirs.shape[0], 3 * dirs.shape[0]).reshape(-1, 3)
        assert_same(got, oracle.texture_cube(env, dirs), f"cubemap N={n}")
        pt.Dispose()


def test_closest_hit_fold_bit_exact(tracer256, oracle, default_scene):
    rng = np.random.default_rng(3)
    n = 200000
    o = (rng.random((n, 3)).astype(np.float32) - np.float32(0.5)) * np.array([40, 25, 25], np.float32) + np.array([0, 0, -10], np.float32)
    o[:20000] = default_scene.spheres[40].Position + (rng.random((20000, 3)).astype(np.float32) - np.float32(0.5)) * np.float32(1.5)
    d = rng.standard_normal((n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    d[:100, 0] = 0   # axis-parallel components: inf slabs
    rays = np.concatenate([o, d], 1).astype(np.float32)
    ref = oracle.ray_trace(rays, default_scene.ubo_bytes(), 256, 48, 7)
    assert ref[:, 2].sum() > 5000
    for op, name in ((4, "packed scene (megakernel view)"), (6, "raw UBO (proxy view)")):
        got = tracer256.DebugEval(op, rays, n, 12 * n).reshape(-1, 12)
        assert_same(got, ref, name)


def test_group_cooperative_fold_bit_exact(ptb, oracle, env256, camera):
    """The tail-of-frame fold (g lanes share one ray, closed-form reduction of the order-dependent closest hit) against the
    sequential oracle fold, for every group size, with many origins INSIDE primitives (the case that needs the second pass)
    and a scene with overlapping boxes/spheres."""
    rng = np.random.default_rng(11)
    for scene in (ptb.load_default_scene(), ptb.synthetic_scene(96, 40, seed=5)):
        pt = make_tracer(ptb, env256, 16, 16, scene, camera)
        n = 6000
        o = (rng.random((n, 3)).astype(np.float32) - np.float32(0.5)) * np.array([40, 25, 25], np.float32) + np.array([0, 0, -10], np.float32)
        centres = np.stack([s.Position for s in scene.spheres])
        pick = rng.integers(0, len(scene.spheres), n // 2)
        radii = np.array([s.Radius for s in scene.spheres], np.float32)[pick, None]
        o[: n // 2] = centres[pick] + (rng.random((n // 2, 3)).astype(np.float32) - np.float32(0.5)) * radii
        d = rng.standard_normal((n, 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
        rays = np.concatenate([o, d], 1).astype(np.float32)
        ref = oracle.ray_trace(rays, scene.ubo_bytes(), scene.max_spheres, len(scene.spheres), len(scene.cuboids))
        assert ref[:, 2].sum() > 500
        for k in (1, 2, 3, 4, 5, 8, 9, 16, 17, 32):
            got = pt.DebugEval(7, np.append(rays.ravel(), np.float32(k)), n, 12 * n).reshape(-1, 12)
            assert_same(got, ref, f"group fold, {k} live rays per warp, {len(scene.spheres)} spheres")
        pt.Dispose()


@pytest.mark.parametrize("mode", [1, 0], ids=["grid", "bvh"])
def test_large_scene_fold_bit_exact(ptb, oracle, env256, camera, mode):
    """Scenes of >= 96 primitives go through a spatial structure in shared memory — a uniform grid walked by a DDA (default) or
    a binary BVH.  Either may only skip primitives that fail the exact test, so the fold must equal the oracle's brute-force
    fold bit for bit — including rays that start inside (overlapping) primitives, axis-parallel rays, rays grazing box faces,
    far-away origins and degenerate / non-finite geometry."""
    rng = np.random.default_rng(17)
    for scene in (ptb.synthetic_scene(1024, 256), ptb.synthetic_scene(200, 60, seed=3)):
        if len(scene.spheres) == 200:      # poison a few primitives: NaN centre, infinite box, inverted box, zero radius
            scene.spheres[5].Position = np.array([np.nan, 0, 0], np.float32)
            scene.spheres[6].Radius = np.float32(0.0)
            scene.cuboids[20].Dimensions = np.array([np.inf, 1, 1], np.float32)
            scene.cuboids[21].Dimensions = np.array([-1.0, 2.0, -0.5], np.float32)
        pt = make_tracer(ptb, env256, 16, 16, scene, camera)
        pt.SetLargeSceneMode(mode)
        n = 60000
        o = (rng.random((n, 3)).astype(np.float32) - np.float32(0.5)) * np.array([44, 28, 28], np.float32) + np.array([0, 0, -10], np.float32)
        centres = np.stack([s.Position for s in scene.spheres])
        pick = rng.integers(0, len(scene.spheres), n // 3)
        radii = np.array([s.Radius for s in scene.spheres], np.float32)[pick, None]
        o[: n // 3] = np.nan_to_num(centres[pick]) + (rng.random((n // 3, 3)).astype(np.float32) - np.float32(0.5)) * radii * np.float32(1.8)
        cpick = rng.integers(0, len(scene.cuboids), n // 6)
        cmin = np.stack([c.Min for c in scene.cuboids])[cpick]; cmax = np.stack([c.Max for c in scene.cuboids])[cpick]
        with np.errstate(all="ignore"):
            inside_box = cmin + rng.random((n // 6, 3)).astype(np.float32) * (cmax - cmin)
        o[n // 3: n // 3 + n // 6] = np.nan_to_num(inside_box, posinf=5.0, neginf=-5.0)
        d = rng.standard_normal((n, 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
        d[-300:-200, 0] = 0; d[-200:-100, 1] = 0; d[-100:, :2] = 0; d[-100:, 2] = 1       # axis-parallel: infinite reciprocals
        o[-50:] *= np.float32(40.0)                                                          # far outside the scene
        o[-1300:-300] = np.round(o[-1300:-300] / np.float32(0.25)) * np.float32(0.25)        # lattice points: cell boundaries, box faces
        rays = np.concatenate([o, d], 1).astype(np.float32)
        ref = oracle.ray_trace(rays, scene.ubo_bytes(), scene.max_spheres, len(scene.spheres), len(scene.cuboids))
        assert ref[:, 2].sum() > 3000 and ref[:, 0].sum() > 30000
        assert pt.SceneInfo(4) == (3 if mode else 1)
        got = pt.DebugEval(11 if mode else 9, rays, n, 12 * n).reshape(-1, 12).copy()
        # column 3 = BVH nodes / grid cells the ray visited: the probe must really have walked the structure (round 1's probe
        # silently took the brute-force fold), and a structure that visits everything culls nothing
        visits = got[:, 3].copy()
        got[:, 3] = 0
        finite = np.isfinite(rays).all(axis=1)
        total = pt.SceneInfo(6) if mode else pt.BvhNodes
        assert total > 0
        if mode == 0:
            assert (visits[finite] >= 1).all(), "a finite ray did not enter the BVH"
        else:
            in_room = finite & (np.abs(rays[:, 0]) < 18) & (np.abs(rays[:, 1]) < 10) & (rays[:, 2] > -20) & (rays[:, 2] < 0)
            assert (visits[in_room] >= 1).all(), "a ray that starts between the primitives did not enter the grid"
        assert visits[finite].mean() < 0.5 * total, f"mean {visits[finite].mean():.1f} of {total} nodes / cells visited: nothing is culled"
        assert_same(got, ref, f"{'grid' if mode else 'BVH'} fold, {len(scene.spheres)} spheres + {len(scene.cuboids)} cuboids")
        with pytest.raises(ptb.PtbError):          # the other structure is not built
            pt.DebugEval(9 if mode else 11, rays[:1], 1, 12)
        pt.Dispose()
    small = make_tracer(ptb, env256, 16, 16, ptb.load_default_scene(), camera)
    with pytest.raises(ptb.PtbError):          # 55 primitives: no large-scene structure to probe
        small.DebugEval(11 if mode else 9, np.zeros(6, np.float32), 1, 12)
    small.Dispose()


@pytest.mark.parametrize("mode", [1, 0], ids=["grid", "bvh"])
def test_bvh_survives_camera_leaving_the_scene(ptb, oracle, env256, mode):
    """The structures' safety margins scale with an extent that includes the camera; moving the camera far away must rebuild them."""
    scene = ptb.synthetic_scene(160, 40, seed=8)
    sc = ptb.scene
    pt = ptb.PathTracer(env256, 96, 54, 8, 1, 20.0, 0.14, max_spheres=scene.max_spheres, max_cuboids=scene.max_cuboids)
    pt.SetLargeSceneMode(mode)
    pt.LoadScene(scene)
    for pos in ([-17.14, 3.53, -8.62], [-900.0, 40.0, 700.0], [3.0e4, 10.0, -2.0e4], [-17.14, 3.53, -8.62]):
        cam = sc.Camera(np.array(pos, np.float32), np.array([0, 1, 0], np.float32), -32.2, 0.8)
        cam_target = sc.Camera(np.array(pos, np.float32), np.array([0, 1, 0], np.float32),
                               float(np.degrees(np.arctan2(-10 - pos[2], -pos[0]))), float(np.degrees(np.arctan2(-pos[1], np.hypot(pos[0], pos[2] + 10)))))
        for c in (cam, cam_target):
            pt.SetCamera(c); pt.ResetRenderer(); pt.Render()
            ref = oracle_render(oracle, sc, scene, c, env256, 96, 54, 1, depth=8)
            assert_same(pt.Result, ref, f"camera at {pos}")
    pt.Dispose()


def test_ray_classification_fold_bit_exact(ptb, oracle, env256, camera):
    """Scenes of <= 64 primitives: RayTrace() runs over the candidates of the ray's class only (origin cell x direction
    bucket -> 64-bit set).  The table may only drop primitives that fail the exact test, so the fold must equal the oracle's
    fold over all primitives bit for bit — for rays starting inside (overlapping) primitives, on cell boundaries, axis-parallel
    rays, rays from outside the grid (full mask), non-finite rays, and degenerate / non-finite / inverted geometry."""
    rng = np.random.default_rng(23)
    poisoned = ptb.synthetic_scene(40, 24, seed=9)
    poisoned.spheres[5].Position = np.array([np.nan, 0, 0], np.float32)
    poisoned.spheres[6].Radius = np.float32(0.0)
    poisoned.spheres[7].Radius = np.float32(-0.7)
    poisoned.cuboids[10].Dimensions = np.array([np.inf, 1, 1], np.float32)
    poisoned.cuboids[11].Dimensions = np.array([-1.0, 2.0, -0.5], np.float32)
    for scene, cells, buckets in ((ptb.load_default_scene(), 13, 12), (ptb.load_default_scene(), 5, 3), (poisoned, 13, 12),
                                  (ptb.synthetic_scene(2, 1, seed=2), 9, 16)):
        pt = make_tracer(ptb, env256, 16, 16, scene, camera)
        pt.SetRayClassification(1, cells, buckets)
        n = 120000
        o = (rng.random((n, 3)).astype(np.float32) - np.float32(0.5)) * np.array([44, 36, 28], np.float32) + np.array([0, 2, -10], np.float32)
        centres = np.nan_to_num(np.stack([s.Position for s in scene.spheres]))
        pick = rng.integers(0, len(scene.spheres), n // 3)
        radii = np.abs(np.array([s.Radius for s in scene.spheres], np.float32))[pick, None]
        o[: n // 3] = centres[pick] + (rng.random((n // 3, 3)).astype(np.float32) - np.float32(0.5)) * radii * np.float32(2.2)
        cpick = rng.integers(0, len(scene.cuboids), n // 6)
        cmin = np.stack([c.Min for c in scene.cuboids])[cpick]; cmax = np.stack([c.Max for c in scene.cuboids])[cpick]
        with np.errstate(all="ignore"):
            inside_box = cmin + rng.random((n // 6, 3)).astype(np.float32) * (cmax - cmin)
        o[n // 3: n // 3 + n // 6] = np.nan_to_num(inside_box, posinf=5.0, neginf=-5.0)
        d = rng.standard_normal((n, 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
        d[-600:-400, 0] = 0; d[-400:-200, 1] = 0; d[-200:-100, :2] = 0; d[-200:-100, 2] = 1       # axis-parallel: infinite reciprocals
        d[-700:-600, 1] = d[-700:-600, 0]                                                        # |dx| == |dy|: major-axis ties
        d[-700:-600] /= np.linalg.norm(d[-700:-600], axis=1, keepdims=True).astype(np.float32)
        o[-100:-50] *= np.float32(40.0)                                                          # far outside the grid
        o[-50:-40] = np.nan; d[-40:-30] = np.nan; d[-30:-20] = 0; o[-20:-10, 0] = np.inf          # non-finite / zero rays
        o[-1000:-700] = np.round(o[-1000:-700] / np.float32(0.5)) * np.float32(0.5)              # lattice points: cell boundaries
        rays = np.concatenate([o, d], 1).astype(np.float32)
        ref = oracle.ray_trace(rays, scene.ubo_bytes(), scene.max_spheres, len(scene.spheres), len(scene.cuboids))
        if len(scene.spheres) + len(scene.cuboids) < 4:
            with pytest.raises(ptb.PtbError):      # too small to be worth a table: plain fold, nothing to probe
                pt.DebugEval(10, rays[:1], 1, 12)
            pt.Dispose()
            continue
        assert pt.SceneInfo(4) == 2 and pt.SceneInfo(5) > 0
        got = pt.DebugEval(10, rays, n, 12 * n).reshape(-1, 12).copy()
        cand = got[:, 3].copy()
        got[:, 3] = 0
        assert_same(got, ref, f"ray-classification fold, {len(scene.spheres)} spheres + {len(scene.cuboids)} cuboids, {cells} cells, {buckets} buckets")
        finite = np.isfinite(rays).all(axis=1) & (np.abs(rays[:, 3:]).sum(axis=1) > 0)
        assert (cand[~finite] == 65).all(), "a non-finite or zero ray must take the full mask"
        assert (cand[-100:-50] == 65).all(), "rays from outside the grid must take the full mask"
        n_prims = len(scene.spheres) + len(scene.cuboids)
        in_grid = finite & (cand < 65)
        assert in_grid.mean() > 0.5
        if cells >= 9:
            assert cand[in_grid].mean() < 0.35 * n_prims, f"mean {cand[in_grid].mean():.1f} candidates of {n_prims}: the table culls nothing"
        hits_per_ray = ref[in_grid, 0].mean()
        assert cand[in_grid].mean() >= hits_per_ray
        pt.Dispose()


def test_ray_classification_does_not_change_the_image(ptb, oracle, env256, default_scene, camera):
    """Table on / off / coarse / fine, geometry edits (rebuild) and material edits (no rebuild): always the oracle's image."""
    sc = ptb.scene
    W, H = 192, 108
    ref = oracle_render(oracle, sc, default_scene, camera, env256, W, H, 3, spp=2)
    for mode, cells, buckets in ((0, 13, 12), (1, 13, 12), (1, 1, 1), (1, 32, 4), (1, 6, 24)):
        pt = make_tracer(ptb, env256, W, H, default_scene, camera, spp=2)
        pt.SetRayClassification(mode, cells, buckets)
        pt.Render(3)
        assert pt.SceneInfo(4) == (2 if mode else 0)
        assert_same(pt.Result, ref, f"ray classification mode {mode}, {cells} cells, {buckets} buckets")
        pt.Dispose()
    # camera outside the room (every primary ray takes the full mask) and inside a glass sphere
    for pos in ([-60.0, 30.0, 40.0], list(default_scene.spheres[40].Position + np.float32(0.2))):
        cam = sc.Camera(np.array(pos, np.float32), np.array([0, 1, 0], np.float32), -32.2, 0.8)
        pt = make_tracer(ptb, env256, W, H, default_scene, cam)
        pt.Render(2)
        assert_same(pt.Result, oracle_render(oracle, sc, default_scene, cam, env256, W, H, 2), f"camera at {pos}")
        pt.Dispose()
    # edits: move a sphere / resize a box (table rebuilt), then change a material only (table kept)
    import copy
    scene = copy.deepcopy(default_scene)
    pt = make_tracer(ptb, env256, W, H, scene, camera)
    pt.Render(1)
    launches = pt.KernelLaunches
    scene.spheres[14].Position = scene.spheres[14].Position + np.array([1.5, -2.0, 3.0], np.float32)
    scene.spheres[14].Upload(pt.GameObjectsUBO)
    scene.cuboids[6].Dimensions = scene.cuboids[6].Dimensions * np.float32(1.7)
    scene.cuboids[6].Upload(pt.GameObjectsUBO)
    pt.ResetRenderer(); pt.Render(2)
    assert_same(pt.Result, oracle_render(oracle, sc, scene, camera, env256, W, H, 2), "after geometry edits")
    rebuilt = pt.KernelLaunches - launches
    scene.spheres[3].Material.Albedo = np.array([0.9, 0.1, 0.2], np.float32)
    scene.spheres[3].Upload(pt.GameObjectsUBO)
    launches = pt.KernelLaunches
    pt.ResetRenderer(); pt.Render(2)
    assert_same(pt.Result, oracle_render(oracle, sc, scene, camera, env256, W, H, 2), "after a material edit")
    assert pt.KernelLaunches - launches == rebuilt - 1, "a material edit must repack the scene but not rebuild the table"
    pt.Dispose()


def test_bvh_survives_camera_leaving_the_scene(ptb, oracle, env256):
    """The BVH's safety margins scale with an extent that includes the camera; moving the camera far away must rebuild them."""
    scene = ptb.synthetic_scene(160, 40, seed=8)
    sc = ptb.scene
    pt = ptb.PathTracer(env256, 96, 54, 8, 1, 20.0, 0.14, max_spheres=scene.max_spheres, max_cuboids=scene.max_cuboids)
    pt.LoadScene(scene)
    for pos in ([-17.14, 3.53, -8.62], [-900.0, 40.0, 700.0], [3.0e4, 10.0, -2.0e4], [-17.14, 3.53, -8.62]):
        cam = sc.Camera(np.array(pos, np.float32), np.array([0, 1, 0], np.float32), -32.2, 0.8)
        cam_target = sc.Camera(np.array(pos, np.float32), np.array([0, 1, 0], np.float32),
                               float(np.degrees(np.arctan2(-10 - pos[2], -pos[0]))), float(np.degrees(np.arctan2(-pos[1], np.hypot(pos[0], pos[2] + 10)))))
        for c in (cam, cam_target):
            pt.SetCamera(c); pt.ResetRenderer(); pt.Render()
            ref = oracle_render(oracle, sc, scene, c, env256, 96, 54, 1, depth=8)
            assert_same(pt.Result, ref, f"camera at {pos}")
    pt.Dispose()


# ------------------------------------------------------------------------------- image parity
CASES = [
    # name, W, H, frames, kwargs
    ("C1 default 256x256", 256, 256, 8, {}),
    ("ragged size (not a multiple of 8)", 251, 123, 3, {}),
    ("SPP 4 (one RNG stream across samples)", 128, 128, 3, dict(spp=4)),
    ("rayDepth 1", 128, 128, 2, dict(depth=1)),
    ("rayDepth 50 (GUI maximum)", 96, 96, 2, dict(depth=50)),
    ("wide aperture, short focal length", 160, 120, 2, dict(focal=5.0, aperture=0.5)),
    ("pinhole (aperture 0)", 160, 120, 2, dict(aperture=0.0)),
    ("1x1 image", 1, 1, 2, {}),
]


@pytest.mark.parametrize("name,W,H,frames,kw", CASES, ids=[c[0] for c in CASES])
def test_image_parity(ptb, oracle, env256, default_scene, camera, name, W, H, frames, kw):
    pt = make_tracer(ptb, env256, W, H, default_scene, camera, **kw)
    ref = np.zeros((H, W, 4), np.float32)
    for f in range(frames):
        pt.Render()
        oracle_render(oracle, ptb.scene, default_scene, camera, env256, W, H, 1, first_frame=f, image=ref, **kw)
        assert_same(pt.Result, ref, f"{name}, frame {f}")
    assert pt.Samples == frames * kw.get("spp", 1) and pt.Frame == frames
    pt.Dispose()


@pytest.mark.parametrize("overlap", [0, 1, 2, 3, 4])
def test_frame_pipelining_modes(ptb, oracle, env256, camera, overlap):
    """ptb_set_overlap: frames traced on alternating streams into scratch images + stream-ordered blend must give the same
    bits as in-place accumulation, also when reads, scene edits and resets are interleaved with the frames in flight."""
    scene = ptb.load_default_scene()
    W, H = 288, 162
    pt = make_tracer(ptb, env256, W, H, scene, camera)
    pt.SetOverlap(overlap)
    ref = np.zeros((H, W, 4), np.float32)
    frame = 0

    def advance(n):
        nonlocal frame
        pt.Render(n)
        oracle_render(oracle, ptb.scene, scene, camera, env256, W, H, n, first_frame=frame, image=ref)
        frame += n

    advance(5)
    assert_same(pt.Result, ref, f"overlap {overlap}: 5 frames back to back")
    for _ in range(3):
        advance(1)
        assert_same(pt.Result, ref, f"overlap {overlap}: read after every frame")
    scene.spheres[7].Material.Emissiv = np.array([3.0, 0.5, 0.2], np.float32)      # scene edit while the pipeline is warm
    scene.spheres[7].Upload(pt.GameObjectsUBO)
    pt.ResetRenderer(); frame = 0
    advance(4)
    assert_same(pt.Result, ref, f"overlap {overlap}: after a scene edit + reset")
    pt.SetKernel(ptb.KERNEL_NAIVE); advance(2); pt.SetKernel(ptb.KERNEL_MEGA); advance(3)
    assert_same(pt.Result, ref, f"overlap {overlap}: proxy and megakernel frames mixed")
    pt.Dispose()


@pytest.mark.parametrize("divisor,overlap", [(2, 2), (3, 4), (8, 3)])
def test_grid_divisor_does_not_change_the_image(ptb, oracle, env256, default_scene, camera, divisor, overlap):
    """ptb_set_grid_divisor: each frame's persistent grid takes a fraction of the CTA slots (several frames co-resident);
    the work counter hands out the same pixels, so the image is the same bits."""
    W, H = 200, 120
    pt = make_tracer(ptb, env256, W, H, default_scene, camera)
    pt.SetOverlap(overlap)
    pt.SetGridDivisor(divisor)
    pt.Render(6)
    ref = oracle_render(oracle, ptb.scene, default_scene, camera, env256, W, H, 6)
    assert_same(pt.Result, ref, f"grid / {divisor}, {overlap} frames in flight")
    with pytest.raises(ptb.PtbError):
        pt.SetGridDivisor(0)
    pt.Dispose()


@pytest.mark.parametrize("batch", [2, 3, 8, 16])
def test_batched_frames_equal_single_frames(ptb, oracle, env256, default_scene, camera, batch):
    """ptb_set_batch: up to `batch` frames per megakernel launch, per-frame blends in order — the same bits as frame by frame,
    also when batches, single frames, resets and read-backs are interleaved and when n is not a multiple of the batch."""
    W, H = 200, 120
    pt = make_tracer(ptb, env256, W, H, default_scene, camera, spp=2)
    pt.SetBatch(batch)
    ref = np.zeros((H, W, 4), np.float32)
    frame = 0

    def advance(n):
        nonlocal frame
        pt.Render(n)
        oracle_render(oracle, ptb.scene, default_scene, camera, env256, W, H, n, first_frame=frame, image=ref, spp=2)
        frame += n

    advance(11)
    assert_same(pt.Result, ref, f"batch {batch}: 11 frames in one call")
    advance(1); advance(batch); advance(1)
    assert_same(pt.Result, ref, f"batch {batch}: single frames around a full batch")
    pt.ResetRenderer(); frame = 0
    advance(2 * batch + 1)
    assert_same(pt.Result, ref, f"batch {batch}: after a reset")
    assert pt.Frame == 2 * batch + 1
    pt.Dispose()


def test_batched_frames_bvh_scene_and_tiles(ptb, oracle, env256, camera):
    sc = ptb.scene
    scene = sc.synthetic_scene(256, 64, seed=5)             # 320 primitives: the BVH instantiation
    W, H = 160, 90
    pt = make_tracer(ptb, env256, W, H, scene, camera, depth=8)
    pt.SetBatch(4)
    pt.Render(6)
    assert_same(pt.Result, oracle_render(oracle, sc, scene, camera, env256, W, H, 6, depth=8), "batch 4, BVH scene")
    pt.Dispose()
    from importlib import import_module
    D = import_module("opentk-pathtracer_b200.distributed")
    default = sc.load_default_scene()
    full = oracle_render(oracle, sc, default, camera, env256, W, H, 5)
    for rank in range(3):
        pt = make_tracer(ptb, env256, W, H, default, camera)
        pt.SetTile(rank, 3, 8)
        pt.SetBatch(4)
        pt.Render(5)
        assert_same(pt.Result, full[D.local_rows_of(rank, 3, 8, H)], f"batch 4, stripes of rank {rank} of 3")
        pt.Dispose()


def test_progressive_accumulation_to_1024_spp(ptb, oracle, env256, default_scene, camera):
    """BASELINE config 2's protocol (frames 0..1023 at SPP 1, running mean) at reduced size: after 1024 pipelined frames the
    accumulation image still equals the oracle's bit for bit — per-channel MSE exactly 0 (the north star asks for < 1e-6)."""
    W, H = 96, 54
    pt = make_tracer(ptb, env256, W, H, default_scene, camera)
    pt.Render(1024)
    got = pt.Result
    ref = oracle_render(oracle, ptb.scene, default_scene, camera, env256, W, H, 1024)
    assert pt.Samples == 1024
    assert_same(got, ref, "1024-spp accumulation")
    assert float(((got[..., :3].astype(np.float64) - ref[..., :3]) ** 2).mean()) == 0.0
    assert np.isfinite(got).all()
    pt.Dispose()


def test_fast_precision_meets_the_north_star_tolerance(ptb, env256, default_scene, camera):
    """SURVEY 8c protocol P2: the fast build (MUFU rcp/rsq/sin/cos/ex2 + FMA contraction, ptb_set_precision) against the exact
    build on BASELINE config 2 itself — default scene, 1920x1080, frames 0..1023 at SPP 1, matched seeds.  Tolerance (north
    star): per-channel MSE < 1e-6 on the accumulated linear image.  Also reported: how many pixels of frame 0 took another
    path (differ by more than rounding), and that the fast build is deterministic."""
    W, H, FRAMES = 1920, 1080, 1024
    images, first = {}, {}
    for prec in (ptb.PRECISION_EXACT, ptb.PRECISION_FAST):
        pt = make_tracer(ptb, env256, W, H, default_scene, camera)
        pt.SetPrecision(prec)
        assert pt.Precision == prec
        pt.Render(1)
        first[prec] = pt.Result[..., :3].astype(np.float64)
        pt.Render(FRAMES - 1)
        images[prec] = pt.Result[..., :3].astype(np.float64)
        if prec == ptb.PRECISION_FAST:
            pt.ResetRenderer(); pt.Render(1)
            assert (pt.Result[..., :3] == first[prec].astype(np.float32)).all(), "the fast build is not deterministic"
            with pytest.raises(ptb.PtbError):
                pt.SetStats(True)                  # statistics are the exact build's
        pt.Dispose()
    a, b = images[ptb.PRECISION_EXACT], images[ptb.PRECISION_FAST]
    assert np.isfinite(a).all() and np.isfinite(b).all()
    mse = ((a - b) ** 2).mean(axis=(0, 1))
    d0 = np.abs(first[ptb.PRECISION_EXACT] - first[ptb.PRECISION_FAST])
    rel = d0 / np.maximum(np.abs(first[ptb.PRECISION_EXACT]), 1e-3)
    diverged = (rel > 1e-3).any(axis=2).mean()
    touched = (d0 > 0).any(axis=2).mean()
    print(f"fast vs exact after {FRAMES} frames: per-channel MSE {mse}, frame 0: {touched * 100:.1f} % of pixels differ in some bit, "
          f"{diverged * 100:.3f} % took another path; mean radiance {a.mean():.4f} vs {b.mean():.4f}")
    assert (mse < 1e-6).all(), f"per-channel MSE {mse} exceeds the north star's 1e-6"
    assert diverged < 0.02
    assert abs(a.mean() - b.mean()) < 1e-3 * a.mean()


def test_fast_precision_other_paths(ptb, oracle, env256, camera):
    """The fast build through the other instantiations: BVH scene, plain fold (table off), SPP > 1, batches, tiles — each
    close to the exact build (same seeds; a few paths per thousand may differ), never NaN where the exact build is finite."""
    sc = ptb.scene
    cases = [("BVH scene", sc.synthetic_scene(256, 64, seed=5), dict(depth=8), {}),
             ("plain fold", sc.load_default_scene(), {}, dict(rct=0)),
             ("SPP 3, batch 4", sc.load_default_scene(), dict(spp=3), dict(batch=4)),
             ("stripes 1 of 3", sc.load_default_scene(), {}, dict(tile=(1, 3, 8)))]
    W, H = 320, 180
    for name, scene, kw, opt in cases:
        res = {}
        for prec in (ptb.PRECISION_EXACT, ptb.PRECISION_FAST):
            pt = make_tracer(ptb, env256, W, H, scene, camera, **kw)
            pt.SetPrecision(prec)
            if "rct" in opt: pt.SetRayClassification(opt["rct"])
            if "batch" in opt: pt.SetBatch(opt["batch"])
            if "tile" in opt: pt.SetTile(*opt["tile"])
            pt.Render(24)
            res[prec] = pt.Result[..., :3].astype(np.float64)
            pt.Dispose()
        a, b = res[ptb.PRECISION_EXACT], res[ptb.PRECISION_FAST]
        assert np.isfinite(b[np.isfinite(a)]).all(), name
        fin = np.isfinite(a) & np.isfinite(b)
        close = np.abs(a - b) <= 1e-3 * np.maximum(np.abs(a), 1e-2)
        assert close[fin].mean() > 0.97, f"{name}: only {close[fin].mean() * 100:.1f} % of the channels agree to 1e-3"
        assert abs(a[fin].mean() - b[fin].mean()) < 0.02 * a[fin].mean(), name


def test_golden_image(ptb, default_scene, camera):
    env16 = np.load(os.path.join(GOLD, "env16.npy"))
    pt = make_tracer(ptb, env16, 64, 64, default_scene, camera)
    pt.Render()
    assert_same(pt.Result, np.load(os.path.join(GOLD, "c1_64x64_f0.npy")), "golden frame 0")
    pt.Render(3)
    assert_same(pt.Result, np.load(os.path.join(GOLD, "c1_64x64_f0_3.npy")), "golden frames 0..3")
    pt.Dispose()


def test_1080p_crop_parity(ptb, oracle, env256, default_scene, camera):
    """C2 at full size: the oracle renders three 24-row bands (sky, mid, floor); the GPU renders everything."""
    W, H = 1920, 1080
    pt = make_tracer(ptb, env256, W, H, default_scene, camera)
    ref = np.zeros((H, W, 4), np.float32)
    bands = [(0, 24), (528, 552), (1056, 1080)]
    for f in range(2):
        pt.Render()
        for b in bands:
            oracle_render(oracle, ptb.scene, default_scene, camera, env256, W, H, 1, first_frame=f, image=ref, rows=b)
    got = pt.Result
    for b in bands:
        assert_same(got[b[0]:b[1]], ref[b[0]:b[1]], f"1080p rows {b}")
    pt.Dispose()


def test_c2_full_size_to_1024_spp(ptb, oracle, env256, default_scene, camera):
    """BASELINE config 2 itself: 1920x1080, frames 0..1023 at SPP 1, running mean.  The GPU renders the whole progressive
    render (batched launches, the path the bench times); the oracle follows three 8-row bands (sky / spheres / floor) through
    all 1024 frames.  Bit-exact: per-channel MSE exactly 0 on 46 080 pixels x 1024 frames."""
    W, H, FRAMES = 1920, 1080, 1024
    pt = make_tracer(ptb, env256, W, H, default_scene, camera)
    pt.Render(FRAMES)
    got = pt.Result
    assert pt.Samples == FRAMES
    bands = [(40, 48), (536, 544), (1000, 1008)]
    ref = np.zeros((H, W, 4), np.float32)
    for b in bands:
        oracle_render(oracle, ptb.scene, default_scene, camera, env256, W, H, FRAMES, image=ref, rows=b)
    for b in bands:
        assert_same(got[b[0]:b[1]], ref[b[0]:b[1]], f"1080p rows {b} after {FRAMES} frames")
    assert np.isfinite(got).all()
    pt.Dispose()


def test_dof_sweep_parity(ptb, oracle, env256, default_scene, camera):
    """C5: ApertureDiameter x FocalLength grid at reduced size."""
    W, H = 96, 54
    for aperture in (0.0, 0.05, 0.3, 0.5):
        for focal in (1.0, 5.0, 50.0):
            pt = make_tracer(ptb, env256, W, H, default_scene, camera, focal=focal, aperture=aperture)
            pt.Render()
            ref = oracle_render(oracle, ptb.scene, default_scene, camera, env256, W, H, 1, focal=focal, aperture=aperture)
            assert_same(pt.Result, ref, f"aperture {aperture} focal {focal}")
            pt.Dispose()


@pytest.mark.parametrize("mode", [1, 0], ids=["grid", "bvh"])
def test_synthetic_scene_parity(ptb, oracle, env256, camera, mode):
    """C3: 1024 spheres + 256 cuboids, rayDepth 8 (capacities 1024/256 move the cuboid base to 81 920), grid and BVH."""
    syn = ptb.synthetic_scene(1024, 256)
    W, H = 160, 90
    pt = make_tracer(ptb, env256, W, H, syn, camera, depth=8)
    pt.SetLargeSceneMode(mode)
    ref = np.zeros((H, W, 4), np.float32)
    for f in range(2):
        pt.Render()
        oracle_render(oracle, ptb.scene, syn, camera, env256, W, H, 1, first_frame=f, image=ref, depth=8)
        assert_same(pt.Result, ref, f"C3 frame {f}")
    pt.Dispose()


def test_empty_scene_and_zero_depth(ptb, oracle, env256, camera):
    empty = ptb.Scene()
    pt = make_tracer(ptb, env256, 64, 48, empty, camera)
    pt.Render()
    assert_same(pt.Result, oracle_render(oracle, ptb.scene, empty, camera, env256, 64, 48, 1), "empty scene")
    pt.RayDepth = 0
    pt.ResetRenderer()
    pt.Render()
    out = pt.Result
    assert (out[..., :3] == 0).all() and (out[..., 3] == 1).all()
    pt.Dispose()


def test_object_count_below_uploaded(ptb, oracle, env256, default_scene, camera):
    """NumSpheres/NumCuboids smaller than what the UBO holds: only the first n are traced (compute.glsl:231,244)."""
    pt = make_tracer(ptb, env256, 96, 96, default_scene, camera)
    pt.NumSpheres = 10
    pt.NumCuboids = 3
    pt.Render()
    sc = ptb.scene
    ref = np.zeros((96, 96, 4), np.float32)
    oracle.render(ref, sc.basic_data_bytes(camera, 96, 96), default_scene.ubo_bytes(), env256, frame=0, spp=1, ray_depth=13,
                  focal_length=20.0, aperture_diameter=0.14, n_spheres=10, n_cuboids=3)
    assert_same(pt.Result, ref, "partial counts")
    pt.Dispose()


def test_object_edit_through_subdata(ptb, oracle, env256, camera):
    """Gui.cs:212-216: editing one object re-uploads just its 80/96 bytes."""
    scene = ptb.load_default_scene()
    pt = make_tracer(ptb, env256, 96, 96, scene, camera)
    pt.Render()
    scene.spheres[5].Material.Emissiv = np.array([2.0, 1.0, 0.5], np.float32)
    scene.spheres[5].Upload(pt.GameObjectsUBO)
    scene.cuboids[6].Dimensions = np.array([5.0, 4.0, 3.0], np.float32)
    scene.cuboids[6].Upload(pt.GameObjectsUBO)
    pt.ResetRenderer()
    pt.Render()
    assert_same(pt.Result, oracle_render(oracle, ptb.scene, scene, camera, env256, 96, 96, 1), "after SubData edits")
    pt.Dispose()


# ------------------------------------------------------------------------------- schedule / partition invariance
def test_megakernel_equals_proxy_at_1080p(ptb, env256, default_scene, camera):
    """Two different schedules (persistent + refill vs one thread per pixel) must give the same bits at full size."""
    imgs = []
    for kernel in (ptb.KERNEL_MEGA, ptb.KERNEL_NAIVE):
        pt = make_tracer(ptb, env256, 1920, 1080, default_scene, camera, kernel=kernel)
        pt.Render(3)
        imgs.append(pt.Result)
        pt.Dispose()
    assert_same(imgs[0], imgs[1], "megakernel vs proxy")
    assert np.isfinite(imgs[0]).all() and (imgs[0][..., 3] == 1).all()


def test_render_is_deterministic(ptb, env256, default_scene, camera):
    a = make_tracer(ptb, env256, 640, 360, default_scene, camera)
    b = make_tracer(ptb, env256, 640, 360, default_scene, camera)
    a.Render(4); b.Render(2); b.Render(2)
    assert_same(a.Result, b.Result, "two contexts, different call batching")
    a.Dispose(); b.Dispose()


@pytest.mark.parametrize("world,stripe", [(2, 8), (3, 8), (8, 16), (4, 5)])
def test_tile_union_equals_full_render(ptb, env256, default_scene, camera, world, stripe):
    """P3: the union of every rank's stripes (global seeds) is bit-identical to the single-GPU image; the device
    de-interleave reproduces the host one."""
    from importlib import import_module
    D = import_module("opentk-pathtracer_b200.distributed")
    W, H = 320, 203
    full = make_tracer(ptb, env256, W, H, default_scene, camera)
    full.Render(2)
    ref = full.Result
    maxr = D.max_local_rows(world, stripe, H)
    gathered = np.zeros((world, maxr, W, 4), np.float32)
    for r in range(world):
        pt = make_tracer(ptb, env256, W, H, default_scene, camera)
        pt.SetTile(r, world, stripe)
        rows = D.local_rows_of(r, world, stripe, H)
        assert pt.Result.shape[0] == rows.size
        pt.Render(2)
        gathered[r, :rows.size] = pt.Result
        pt.Dispose()
    assert_same(D.deinterleave_host(gathered, H, world, stripe), ref, f"world {world}")
    # device de-interleave through the ABI
    import torch
    g = torch.from_numpy(gathered).cuda()
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    full.SetTile(0, world, stripe)
    rc = full._L.ptb_deinterleave_device(full._ctx, C.c_void_p(g.data_ptr()), C.c_void_p(out.data_ptr()))
    assert rc == 0
    full.Synchronize()
    assert_same(out.cpu().numpy(), ref, "device de-interleave")
    full.Dispose()


@pytest.mark.parametrize("world,stripe,fmt", [(2, 8, 1), (3, 8, 0), (8, 16, 2), (4, 5, 1), (1, 8, 1)])
def test_scatter_readback_assembles_the_frame_on_the_host(ptb, oracle, env256, default_scene, camera, world, stripe, fmt):
    """ptb_read_result_scatter_async: every rank copies its stripes into their rows of ONE full-frame pinned host image
    (ragged last stripe included, H = 203); the assembled frame equals the single-GPU image in all three formats."""
    import torch
    W, H = 320, 203
    full = make_tracer(ptb, env256, W, H, default_scene, camera)
    full.Render(2)
    ref = full.Result
    full.Dispose()
    want = {0: ref, 1: ref[..., :3], 2: oracle.tonemap(ref)}[fmt]
    host = torch.zeros(want.shape, dtype=torch.uint8 if fmt == 2 else torch.float32).pin_memory()
    for r in range(world):
        pt = make_tracer(ptb, env256, W, H, default_scene, camera)
        pt.SetTile(r, world, stripe)
        pt.Render(2)
        pt.ReadResultScatterAsync(host.data_ptr(), fmt)
        pt.Synchronize()
        pt.Dispose()
    if fmt == 2:
        assert (host.numpy() == want).all()
    else:
        assert_same(host.numpy(), want, f"world {world}, stripe {stripe}, format {fmt}")


# ------------------------------------------------------------------------------- PathTracer class semantics
def test_reset_setsize_and_resume(ptb, oracle, env256, default_scene, camera):
    pt = make_tracer(ptb, env256, 128, 96, default_scene, camera)
    pt.Render(3)
    assert pt.Samples == 3
    pt.ResetRenderer()                       # PathTracer.cs:137-140: counter only; frame 0 then overwrites the image
    assert pt.Samples == 0
    pt.Render()
    f0 = oracle_render(oracle, ptb.scene, default_scene, camera, env256, 128, 96, 1)
    assert_same(pt.Result, f0, "frame 0 after ResetRenderer")
    # checkpoint / resume: save image + frame, restore into a new context, continue
    pt.Render(2)
    saved, frame = pt.Result, pt.Frame
    other = make_tracer(ptb, env256, 128, 96, default_scene, camera)
    other.WriteResult(saved); other.SetFrame(frame)
    pt.Render(2); other.Render(2)
    assert_same(pt.Result, other.Result, "resumed accumulation")
    other.Dispose()
    pt.SetSize(64, 80)                       # PathTracer.cs:131-135
    assert pt.Samples == 0 and pt.Width == 64 and pt.Height == 80
    pt.SetCamera(camera)                     # MainWindow.OnResize re-uploads the projection (aspect changed)
    pt.Render()
    assert_same(pt.Result, oracle_render(oracle, ptb.scene, default_scene, camera, env256, 64, 80, 1), "after SetSize")
    pt.Dispose()


def test_pipelined_readback(ptb, env256, default_scene, camera):
    """ptb_read_result_async: snapshot + copy on a second stream; two reads in flight while the next frames render."""
    import torch
    pt = make_tracer(ptb, env256, 320, 200, default_scene, camera)
    bufs = [torch.empty((200, 320, 4), dtype=torch.float32).pin_memory() for _ in range(3)]
    for f in range(3):
        pt.Render()
        pt.ReadResultAsync(bufs[f].data_ptr())
    pt.Synchronize()
    ref = make_tracer(ptb, env256, 320, 200, default_scene, camera)
    for f in range(3):
        ref.Render()
        assert_same(bufs[f].numpy(), ref.Result, f"async read of frame {f}")
    pt.Dispose(); ref.Dispose()


def test_readback_formats(ptb, oracle, env256, default_scene, camera):
    """ptb_read_result_format_async: RGB32F is the colour floats bit for bit (alpha is the constant 1.0 of compute.glsl:129),
    RGBA8 is ScreenEffect's pass fused into the snapshot; ragged pixel counts exercise the packer's tail."""
    import torch
    for W, H in ((320, 200), (67, 3), (1, 1)):
        pt = make_tracer(ptb, env256, W, H, default_scene, camera)
        rgb = [torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
        rgba = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
        ldr = torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory()
        pt.Render()
        pt.ReadResultAsync(rgb[0].data_ptr(), ptb.FORMAT_RGB32F)
        pt.Render()
        pt.ReadResultAsync(rgb[1].data_ptr(), ptb.FORMAT_RGB32F)
        pt.ReadResultAsync(rgba.data_ptr(), ptb.FORMAT_RGBA32F)
        pt.ReadResultAsync(ldr.data_ptr(), ptb.FORMAT_RGBA8)
        pt.Synchronize()
        full = pt.Result
        assert (full[..., 3] == 1).all()
        assert_same(rgba.numpy(), full, f"{W}x{H} RGBA32F")
        assert_same(rgb[1].numpy(), full[..., :3], f"{W}x{H} RGB32F")
        assert (ldr.numpy() == oracle.tonemap(full)).all(), f"{W}x{H} RGBA8"
        ref = make_tracer(ptb, env256, W, H, default_scene, camera)
        ref.Render()
        assert_same(rgb[0].numpy(), ref.Result[..., :3], f"{W}x{H} RGB32F of the first frame (read while the second rendered)")
        pt.Dispose(); ref.Dispose()
    with pytest.raises(ptb.PtbError):
        pt2 = make_tracer(ptb, env256, 8, 8, default_scene, camera)
        try:
            pt2.ReadResultAsync(rgba.data_ptr(), 7)
        finally:
            pt2.Dispose()


def test_error_codes(ptb, env256):
    L = ptb.load_library()
    ctx = C.c_void_p()
    assert L.ptb_create(C.byref(ctx), 0, 10, 256, 64, 0) == -1
    assert L.ptb_create(C.byref(ctx), 16, 16, 256, 64, 99) == -1 and b"device" in L.ptb_last_error()
    assert L.ptb_create(C.byref(ctx), 16, 16, 256, 64, 0) == 0
    assert L.ptb_render(ctx) == -3 and b"EnvironmentMap" in L.ptb_last_error()      # PTB_E_STATE
    assert L.ptb_set_num_spheres(ctx, 257) == -1 and L.ptb_set_num_cuboids(ctx, -1) == -1
    assert L.ptb_set_spp(ctx, 0) == -1 and L.ptb_set_ray_depth(ctx, -1) == -1
    buf = C.create_string_buffer(256)
    assert L.ptb_basic_data_subdata(ctx, 140, 16, buf) == -1
    assert L.ptb_game_objects_subdata(ctx, 256 * 80 + 64 * 96 - 8, 16, buf) == -1
    assert L.ptb_game_objects_subdata(ctx, 256 * 80 + 64 * 96 - 16, 16, buf) == 0
    assert L.ptb_set_tile(ctx, 2, 2, 8) == -1 and L.ptb_set_kernel(ctx, 7) == -1
    L.ptb_destroy(ctx)
    # a scene whose geometry cannot be staged in shared memory is refused, not silently truncated
    huge = ptb.PathTracer(env256, 16, 16, 4, 1, 20.0, 0.14, max_spheres=16384, max_cuboids=64)
    huge.NumSpheres = 16384
    with pytest.raises(ptb.PtbError):
        huge.Render()
    huge.Dispose()


def test_gl_interop_entry_points_without_a_gl_context(ptb, env256, default_scene, camera):
    """PathTracer.Result is a GL texture in the reference (PathTracer.cs:86,97-99); ptb_register_gl_texture / ptb_present_gl hand
    it over through CUDA-GL interop.  A headless box has no GL context: registration must fail cleanly with PTB_E_STATE (-3),
    leave the context usable, and presenting without a registered texture is a call-order error."""
    pt = make_tracer(ptb, env256, 64, 48, default_scene, camera)
    with pytest.raises(ptb.PtbError, match="error -3"):
        pt.RegisterGLTexture(1)
    with pytest.raises(ptb.PtbError, match="error -3"):
        pt.PresentGL()
    pt.Render(2)
    assert np.isfinite(pt.Result).all()
    pt.Dispose()


def test_four_thousand_spheres(ptb, oracle, env256, camera):
    """4096 spheres + 64 cuboids: 327 KB as a packed block, but with the BVH only geometry + hierarchy (~190 KB) are staged and
    the materials stay in HBM."""
    scene = ptb.synthetic_scene(4096, 64, seed=77)
    pt = make_tracer(ptb, env256, 48, 32, scene, camera, depth=6)
    pt.Render(2)
    ref = oracle_render(oracle, ptb.scene, scene, camera, env256, 48, 32, 2, depth=6)
    assert_same(pt.Result, ref, "4096-sphere scene")
    pt.Dispose()


# ------------------------------------------------------------------------------- atmosphere producer on the GPU
def test_atmosphere_kernel_bit_exact(ptb, oracle, env256, default_scene, camera):
    pt = ptb.PathTracer(None, 32, 32, 13, 1, 20.0, 0.14)
    pt.GenerateAtmosphere(256, 50, 15, 0.5, 15.0)
    assert_same(pt.ReadEnvironment(), env256, "atmosphere 256, 50x15")
    sc = ptb.scene
    pt.GenerateAtmosphere(33, 7, 3, 0.3, 22.0)
    assert_same(pt.ReadEnvironment(), oracle.atmosphere(33, sc.atmosphere_ubo_bytes(), sc.atmosphere_light_pos(0.3), 22.0, 7, 3), "atmosphere 33, 7x3")
    # and it is usable as the environment map
    pt.LoadScene(default_scene); pt.SetCamera(camera)
    pt.Render()
    ref = oracle_render(oracle, sc, default_scene, camera, oracle.atmosphere(33, sc.atmosphere_ubo_bytes(), sc.atmosphere_light_pos(0.3), 22.0, 7, 3), 32, 32, 1)
    assert_same(pt.Result, ref, "render with generated atmosphere")
    pt.Dispose()


def test_fast_atmosphere_against_the_exact_kernel(ptb, env256, default_scene, camera):
    """ptb_generate_atmosphere_fast (the same loops with special-function exp / sqrt / rcp and fused multiply-adds)
    against ptb_generate_atmosphere (bit-exact with the shader): several sun positions, step counts and sizes.  Tolerance: 1e-3
    of the map's brightest texel per channel, mean relative error below 1e-4 (the inputs are ~6.4e6 m: a single-precision
    square root of a squared radius carries ~1 m of noise into exp(-h / 1200 m), in either kernel)."""
    pt = make_tracer(ptb, env256, 32, 32, default_scene, camera)
    for size, i_steps, j_steps, time, intensity in ((64, 50, 15, 0.5, 15.0), (96, 30, 8, 0.3, 15.0), (48, 100, 40, 0.07, 22.0), (40, 20, 10, 0.75, 15.0), (33, 12, 3, 0.62, 5.0)):
        pt.GenerateAtmosphere(size, i_steps, j_steps, time, intensity)
        exact = pt.ReadEnvironment()[..., :3].astype(np.float64)
        pt.GenerateAtmosphere(size, i_steps, j_steps, time, intensity, fast=True)
        fast = pt.ReadEnvironment()[..., :3].astype(np.float64)
        assert np.isfinite(fast).all()
        scale = max(exact.max(), 1e-3)          # (a sun far below the horizon leaves the whole map at ~0)
        err = np.abs(fast - exact)
        assert err.max() <= 1e-3 * scale, f"size {size} time {time}: max error {err.max():.3e} vs brightest texel {scale:.3e}"
        lit = exact > 1e-3 * scale
        if lit.any():
            assert (err[lit] / exact[lit]).mean() < 1e-4, f"size {size} time {time}: mean relative error {(err[lit] / exact[lit]).mean():.2e}"
    # and it renders: same image statistics through either environment
    pt.GenerateAtmosphere(64, 50, 15, 0.5, 15.0); pt.ResetRenderer(); pt.Render(8); a = pt.Result[..., :3].mean()
    pt.GenerateAtmosphere(64, 50, 15, 0.5, 15.0, fast=True); pt.ResetRenderer(); pt.Render(8); b = pt.Result[..., :3].mean()
    assert abs(a - b) < 1e-3 * a
    pt.Dispose()


def test_statistics_counters(ptb, oracle, env256, default_scene, camera):
    pt = make_tracer(ptb, env256, 128, 128, default_scene, camera)
    pt.SetStats(True)
    pt.Render()
    st = pt.ReadStats()
    sc = ptb.scene
    img = np.zeros((128, 128, 4), np.float32)
    ref = oracle.render(img, sc.basic_data_bytes(camera, 128, 128), default_scene.ubo_bytes(), env256, frame=0, spp=1, ray_depth=13,
                        focal_length=20.0, aperture_diameter=0.14, n_spheres=48, n_cuboids=7, want_stats=True)
    assert st["samples"] == ref["samples"] and st["bounces"] == ref["bounces"] and st["hits"] == ref["hits"]
    pt.Dispose()


# ------------------------------------------------------------------------------- next rows: N1 post-process, N2 sRGB skybox
def test_log_bit_exact(tracer256, oracle):
    x = np.concatenate([np.logspace(-44, 38, 400001).astype(np.float32), np.linspace(0.25, 4.0, 400001, dtype=np.float32),
                        np.array([0.0, -0.0, 1.0, np.inf, -1.0, np.nan, 1e-45], np.float32)])
    assert_same(tracer256.DebugEval(8, x, x.size, x.size), oracle.log(x), "log")


def test_tonemap_bit_exact(ptb, oracle, env256, default_scene, camera):
    """ScreenEffect.Render(PathTracer.Result) on the GPU == the oracle's pass over the same float image, byte for byte."""
    pt = make_tracer(ptb, env256, 320, 180, default_scene, camera)
    pt.Render(4)
    fx = ptb.ScreenEffect()
    got = fx.Render(pt)
    assert got.shape == (180, 320, 4) and got.dtype == np.uint8
    want = oracle.tonemap(pt.Result)
    assert (got == want).all(), f"{int((got != want).any(axis=2).sum())} pixels differ"
    assert 5 < got[..., :3].mean() < 250 and (got[..., 3] == 255).all()
    # a synthetic HDR ramp incl. the linear toe, values > 1 and garbage
    ramp = np.zeros((180, 320, 4), np.float32)
    ramp[..., :3] = np.linspace(0, 12, 180 * 320 * 3, dtype=np.float32).reshape(180, 320, 3) ** 2 / 12
    ramp[0, :4, 0] = [np.nan, -3.0, np.inf, 1e-8]
    pt.WriteResult(ramp)
    assert (fx.Render(pt) == oracle.tonemap(ramp)).all()
    pt.Dispose()


def test_srgb8_skybox_environment(ptb, oracle, default_scene, camera):
    """EnvironmentMap = SkyBox (Srgb8Alpha8 faces, Helper.cs:18-50): decode on upload == oracle decode, and the render
    with it == the oracle render with the decoded float cubemap."""
    rng = np.random.default_rng(21)
    faces = rng.integers(0, 256, (6, 24, 24, 4), dtype=np.uint8)
    pt = ptb.PathTracer(None, 96, 64, 13, 1, 20.0, 0.14)
    pt.SetSkyBox(faces)
    lin = oracle.srgb8_to_linear(faces)
    assert_same(pt.ReadEnvironment(), lin, "sRGB8 -> linear decode")
    pt.LoadScene(default_scene); pt.SetCamera(camera)
    pt.Render(2)
    ref = oracle_render(oracle, ptb.scene, default_scene, camera, lin, 96, 64, 2)
    assert_same(pt.Result, ref, "render with the sRGB skybox")
    pt.Dispose()


# ------------------------------------------------------------------------------- outputs of the reference's own shaders
# tests/golden/ref_*.npz come from the reference's GLSL compiled for the CPU (oracle/build_ref.py; generator
# tests/golden/make_ref_golden.py).  The CUDA path is driven with the stored input BYTES through the two SubData entry
# points, exactly as the C# host would, and must reproduce the stored images bit for bit.
def _tracer_from_bytes(ptb, env, W, H, basic, ubo, ns, nc, max_spheres=256, max_cuboids=64, **kw):
    pt = ptb.PathTracer(env, W, H, kw["depth"], kw["spp"], kw["focal"], kw["aperture"], max_spheres=max_spheres, max_cuboids=max_cuboids)
    pt.BasicDataUBO.SubData(0, len(basic), basic)
    pt.GameObjectsUBO.SubData(0, len(ubo), ubo)
    pt.NumSpheres, pt.NumCuboids = ns, nc
    return pt


@pytest.mark.parametrize("kernel", [0, 1], ids=["megakernel", "gl-proxy"])
def test_reference_golden_default_scene(ptb, kernel):
    g = np.load(os.path.join(GOLD, "ref_pt_default.npz"))
    n, H, W, _ = g["after_frame"].shape
    pt = _tracer_from_bytes(ptb, g["env"], W, H, g["basic_ubo"].tobytes(), g["objects_ubo"].tobytes(), 48, 7,
                            depth=13, spp=2, focal=20.0, aperture=0.14)
    pt.SetKernel(kernel)
    for f in range(n):
        pt.Render()
        assert_same(pt.Result, g["after_frame"][f], f"reference golden, default scene, frame {f}")
    pt.Dispose()


def test_reference_golden_synthetic_scene(ptb):
    g = np.load(os.path.join(GOLD, "ref_pt_synthetic.npz"))
    env = np.load(os.path.join(GOLD, "ref_pt_default.npz"))["env"]
    n, H, W, _ = g["after_frame"].shape
    pt = _tracer_from_bytes(ptb, env, W, H, g["basic_ubo"].tobytes(), g["objects_ubo"].tobytes(), 256, 64,
                            depth=8, spp=1, focal=8.0, aperture=0.4)            # 320 primitives: the BVH fold
    pt.WriteResult(np.zeros((H, W, 4), np.float32))
    pt.SetFrame(5)
    for k in range(n):
        pt.Render()
        assert_same(pt.Result, g["after_frame"][k], f"reference golden, synthetic scene, frame {5 + k}")
    pt.Dispose()


def test_reference_golden_config3_scene(ptb):
    """BASELINE config 3 (1024 spheres + 256 cuboids, capacities 1024 / 256: BVH fold, materials left in HBM) against the
    reference shader built with its two UBO array lengths rewritten (build_ref.py --capacity 1024 256)."""
    g = np.load(os.path.join(GOLD, "ref_pt_config3.npz"))
    env = np.load(os.path.join(GOLD, "ref_pt_default.npz"))["env"]
    n, H, W, _ = g["after_frame"].shape
    pt = _tracer_from_bytes(ptb, env, W, H, g["basic_ubo"].tobytes(), g["objects_ubo"].tobytes(), 1024, 256,
                            max_spheres=1024, max_cuboids=256, depth=8, spp=1, focal=20.0, aperture=0.14)
    for f in range(n):
        pt.Render()
        assert_same(pt.Result, g["after_frame"][f], f"reference golden, config 3, frame {f}")
    pt.Dispose()


def test_reference_golden_atmosphere_and_post(ptb):
    g = np.load(os.path.join(GOLD, "ref_atmosphere.npz"))
    pt = ptb.PathTracer(None, 32, 24, 13, 1, 20.0, 0.14)
    pt.GenerateAtmosphere(16, 8, 4, 0.5, 15.0)
    assert_same(pt.ReadEnvironment(), g["faces_a"], "reference golden, atmosphere 16")
    pt.GenerateAtmosphere(12, 6, 3, 0.2, 22.0)
    assert_same(pt.ReadEnvironment(), g["faces_b"], "reference golden, atmosphere 12")
    p = np.load(os.path.join(GOLD, "ref_post.npz"))
    pt.WriteResult(p["ramp"])
    assert (ptb.ScreenEffect().Render(pt) == p["ramp_rgba8"]).all()
    final = np.load(os.path.join(GOLD, "ref_pt_default.npz"))["after_frame"][-1]
    pt.SetSize(final.shape[1], final.shape[0])
    pt.WriteResult(final)
    assert (ptb.ScreenEffect().Render(pt) == p["rendered"]).all()
    pt.Dispose()


# ------------------------------------------------------------------------------- the reference's shader, compiled for the GPU
def test_reference_shader_compiled_by_nvcc_equals_the_megakernel(ptb):
    """The GL-compute proxy built from the reference's own source (build_ref.py --cuda), dispatched in the reference's launch
    shape, against the product's megakernel on the stored golden inputs: bit for bit."""
    from oracle import ref_cuda
    g = np.load(os.path.join(GOLD, "ref_pt_default.npz"))
    n, H, W, _ = g["after_frame"].shape
    proxy = ref_cuda.CudaReference(fast=False)
    img = np.zeros((H, W, 4), np.float32)
    for f in range(n):
        proxy.render(img, g["basic_ubo"].tobytes(), g["objects_ubo"].tobytes(), g["env"], frame=f, frames=1, spp=2, ray_depth=13,
                     focal_length=20.0, aperture_diameter=0.14, n_spheres=48, n_cuboids=7)
        assert_same(img, g["after_frame"][f], f"nvcc build of compute.glsl, frame {f}")


@pytest.mark.parametrize("seed", range(4))
def test_randomised_configurations_on_the_gpu(ptb, oracle, seed):
    """The CPU pin's forty random dispatches (tests/test_reference_pin.py::test_randomised_configurations), replayed through the
    C ABI: camera inside boxes and spheres, axis-aligned views, any FOV / image shape / SPP / rayDepth / lens / frame number,
    partial object counts, raw material bytes beyond the C# clamps — megakernel (brute force or BVH) against the oracle."""
    sc = ptb.scene
    rng = np.random.default_rng(1000 + seed)
    env16 = oracle.atmosphere(16, sc.atmosphere_ubo_bytes(), sc.atmosphere_light_pos(0.5), 15.0, 8, 4)
    base = sc.synthetic_scene(256, 64, seed=50 + seed)
    raw0 = np.frombuffer(base.ubo_bytes(), np.float32).copy()
    for trial in range(10):
        raw = raw0.copy()
        sph = raw[:256 * 20].reshape(256, 20)
        cub = raw[256 * 20:].reshape(64, 24)
        k = rng.integers(0, 256, 24)
        sph[k, 7] = rng.uniform(-0.2, 1.3, k.size)
        sph[k, 15] = rng.uniform(-0.2, 1.3, k.size)
        sph[k, 16] = rng.choice([0.0, 0.3, 1.0], k.size)
        sph[k, 17] = rng.uniform(0.5, 2.5, k.size)
        cub[rng.integers(0, 64, 6), 15 + 4] = rng.uniform(0.0, 1.0, 6)
        W, H = int(rng.integers(1, 70)), int(rng.integers(1, 40))
        cam = sc.default_camera()
        mode = trial % 4
        if mode == 0:
            cam.Position = (np.asarray(base.spheres[int(rng.integers(0, 256))].Position, np.float32)).copy()
        elif mode == 1:
            c = base.cuboids[int(rng.integers(7, 64))]
            cam.Position = ((c.Min + c.Max) * np.float32(0.5)).astype(np.float32)
        elif mode == 2:
            cam.Position = rng.uniform([-19, -11, -21], [19, 11, 1]).astype(np.float32)
        cam.LookX = float(rng.choice([0.0, 90.0, -90.0, 180.0, rng.uniform(-180, 180)]))
        cam.LookY = float(rng.choice([0.0, 89.0, -89.0, rng.uniform(-80, 80)]))
        fov = np.float32(rng.uniform(20, 140))
        basic, ubo = sc.basic_data_bytes(cam, W, H, fov), raw.tobytes()
        ns = float(rng.choice([256, 0, rng.integers(0, 257), rng.uniform(0, 256)]))
        nc = float(rng.choice([64, 0, rng.integers(0, 65)]))
        ns, nc = int(np.ceil(ns)), int(np.ceil(nc))      # the C# property is an int; `i < count` with a fractional count = ceil
        start = rng.random((H, W, 4)).astype(np.float32)
        first = int(rng.choice([0, 1, 2, 77, 4095, 1 << 20]))
        kw = dict(spp=int(rng.integers(1, 4)), ray_depth=int(rng.choice([0, 1, 2, 8, 13, 30])), focal_length=float(rng.uniform(0.5, 60)),
                  aperture_diameter=float(rng.choice([0.0, 0.14, rng.uniform(0, 2)])))
        pt = ptb.PathTracer(env16, W, H, kw["ray_depth"], kw["spp"], kw["focal_length"], kw["aperture_diameter"])
        pt.BasicDataUBO.SubData(0, len(basic), basic)
        pt.GameObjectsUBO.SubData(0, len(ubo), ubo)
        pt.NumSpheres, pt.NumCuboids = ns, nc
        pt.WriteResult(start)
        pt.SetFrame(first)
        ref = start.copy()
        for f in range(first, first + 2):
            pt.Render()
            oracle.render(ref, basic, ubo, env16, frame=f, n_spheres=ns, n_cuboids=nc, **kw)
            assert_same(pt.Result, ref, f"seed {seed} trial {trial} frame {f}: {W}x{H} {kw} counts=({ns},{nc})")
        pt.Dispose()
