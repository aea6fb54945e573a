"""Generates tests/golden/ref_*.npz: OUTPUTS OF THE REFERENCE ITSELF — its three GLSL shaders compiled for the CPU from
/root/reference by oracle/build_ref.py (oracle/_ref/libglsl_ref.so) — together with the exact input bytes they were
produced from.  The oracle restatement (always) and the CUDA path (on the GPU box) are tested against these files, so the
pin survives on machines where neither /root/reference nor oracle/_ref exists.

    python oracle/build_ref.py && python tests/golden/make_ref_golden.py

ref_pt_default.npz    default scene (MainWindow.cs:208-267), 96x54, SPP 2, frames 0..2 accumulated, rayDepth 13, focal 20,
                      aperture 0.14 (MainWindow.cs:190), 32^2 atmosphere (10 x 4 steps); the running mean after each frame.
ref_pt_synthetic.npz  256 spheres + 64 cuboids with random materials (scene.synthetic_scene(256, 64, seed=7)), 64x36, SPP 1,
                      rayDepth 8, frames 5..6 on top of a zero image, wide aperture.
ref_pt_config3.npz    BASELINE config 3's scene (scene.synthetic_scene(1024, 256), capacities 1024 / 256), 96x54, SPP 1, rayDepth 8,
                      frames 0..1 — from the build with the two UBO array lengths rewritten (build_ref.py --capacity 1024 256).
ref_atmosphere.npz    AtmosphericScattering/compute.glsl: 16^2 (8 x 4 steps, time 0.5) and 12^2 (6 x 3 steps, time 0.2).
ref_post.npz          PostProcessing/fragment.glsl over the final ref_pt_default image and over a synthetic HDR ramp.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def hdr_ramp() -> np.ndarray:
    rng = np.random.default_rng(77)
    img = (rng.random((24, 32, 4)).astype(np.float32) * np.float32(10.0) ** rng.integers(-6, 3, (24, 32, 1)).astype(np.float32)).astype(np.float32)
    img[0, :8, 0] = [0.0, -0.0, -1.0, np.nan, np.inf, -np.inf, 1e-30, 0.0031308]
    return img


def generate(R) -> dict:
    """{file name: {array name: array}} computed with the compiled reference shaders `R` (oracle.ref)."""
    import ptb200
    sc = ptb200.scene
    out = {}

    atmo_ubo = sc.atmosphere_ubo_bytes()
    env32 = R.atmosphere(32, atmo_ubo, sc.atmosphere_light_pos(0.5), 15.0, 10, 4)
    out["ref_atmosphere.npz"] = dict(
        ubo=np.frombuffer(atmo_ubo, np.uint8).copy(),
        light_pos_a=np.asarray(sc.atmosphere_light_pos(0.5), np.float32), faces_a=R.atmosphere(16, atmo_ubo, sc.atmosphere_light_pos(0.5), 15.0, 8, 4),
        light_pos_b=np.asarray(sc.atmosphere_light_pos(0.2), np.float32), faces_b=R.atmosphere(12, atmo_ubo, sc.atmosphere_light_pos(0.2), 22.0, 6, 3))

    scene, cam = sc.load_default_scene(), sc.default_camera()
    W, H = 96, 54
    basic, ubo = sc.basic_data_bytes(cam, W, H), scene.ubo_bytes()
    img = np.zeros((H, W, 4), np.float32)
    frames = []
    for f in range(3):
        R.render(img, basic, ubo, env32, frame=f, spp=2, ray_depth=13, focal_length=20.0, aperture_diameter=0.14, n_spheres=48, n_cuboids=7)
        frames.append(img.copy())
    out["ref_pt_default.npz"] = dict(basic_ubo=np.frombuffer(basic, np.uint8).copy(), objects_ubo=np.frombuffer(ubo, np.uint8).copy(),
                                     env=env32, after_frame=np.stack(frames))

    scene = sc.synthetic_scene(256, 64, seed=7)
    W, H = 64, 36
    basic, ubo = sc.basic_data_bytes(cam, W, H), scene.ubo_bytes()
    img2 = np.zeros((H, W, 4), np.float32)
    frames2 = []
    for f in (5, 6):
        R.render(img2, basic, ubo, env32, frame=f, spp=1, ray_depth=8, focal_length=8.0, aperture_diameter=0.4, n_spheres=256, n_cuboids=64)
        frames2.append(img2.copy())
    out["ref_pt_synthetic.npz"] = dict(basic_ubo=np.frombuffer(basic, np.uint8).copy(), objects_ubo=np.frombuffer(ubo, np.uint8).copy(),
                                       after_frame=np.stack(frames2))

    from oracle import build_ref
    lib = build_ref.capacity_lib((1024, 256))
    if not os.path.exists(lib) and os.path.isdir("/root/reference"):
        build_ref.build(capacity=(1024, 256))
    if os.path.exists(lib):                                   # absent only on a machine that never saw /root/reference
        big = R.variant(lib)
        scene = sc.synthetic_scene(1024, 256)
        W, H = 96, 54
        basic, ubo = sc.basic_data_bytes(cam, W, H), scene.ubo_bytes()
        img3 = np.zeros((H, W, 4), np.float32)
        frames3 = []
        for f in (0, 1):
            big.render(img3, basic, ubo, env32, frame=f, spp=1, ray_depth=8, focal_length=20.0, aperture_diameter=0.14, n_spheres=1024,
                       n_cuboids=256, max_spheres=1024)
            frames3.append(img3.copy())
        out["ref_pt_config3.npz"] = dict(basic_ubo=np.frombuffer(basic, np.uint8).copy(), objects_ubo=np.frombuffer(ubo, np.uint8).copy(),
                                         after_frame=np.stack(frames3))

    ramp = hdr_ramp()
    out["ref_post.npz"] = dict(rendered=R.post(img), ramp=ramp, ramp_rgba8=R.post(ramp))
    return out


def main():
    from oracle import ref as R
    if not R.available():
        from oracle import build_ref
        build_ref.build()
    for name, arrays in generate(R).items():
        np.savez_compressed(os.path.join(HERE, name), **arrays)
        print(name, os.path.getsize(os.path.join(HERE, name)), "bytes")


if __name__ == "__main__":
    main()
