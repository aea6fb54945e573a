"""Generates the committed fixtures under tests/golden/.

pcg_kats.json     integer-exact known answers for the seed (compute.glsl:106) and the PCG stream (compute.glsl:334-344),
                  computed HERE with Python integers only (independent of the oracle and of the product);
                  the same six vectors are tabulated in SURVEY.md §8c.
layout_pins.json  byte sizes / offsets the host relies on (Material.cs:9, Sphere.cs:8,20, Cuboid.cs:8,21, MainWindow.cs:196,200).
env16.npy, c1_64x64_f0.npy, c1_64x64_f0_3.npy
                  regression pins of the oracle itself (a 16^2 atmosphere cubemap, and the default scene at 64x64:
                  frame 0, and the running mean after frames 0..3).  Written by the oracle; tests/test_reference_pin.py checks
                  that the compiled reference shaders (oracle/_ref) produce the same files, and the reference's own outputs
                  are the ref_*.npz files written by make_ref_golden.py.
Run: python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

M32 = 0xFFFFFFFF


def seed(x, y, frame):
    return ((x * 1973 + y * 9277 + frame * 2699) & M32) | 1


def pcg(state):
    state = (state * 747796405 + 2891336453) & M32
    word = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & M32
    return state, ((word >> 22) ^ word) & M32


def main():
    kats = []
    for (x, y, f) in [(0, 0, 0), (1, 0, 0), (0, 1, 0), (5, 7, 3), (1919, 1079, 0), (3839, 2159, 1023), (251, 122, 77), (4095, 4095, 100000)]:
        s = seed(x, y, f)
        st, hs = s, []
        for _ in range(8):
            st, h = pcg(st)
            hs.append(h)
        kats.append(dict(x=x, y=y, frame=f, seed=s, hashes=[f"{h:08x}" for h in hs],
                         floats_hex=[np.float32(np.float32(h) * np.float32(2.0 ** -32)).view(np.uint32).item() for h in hs]))
    json.dump(kats, open(os.path.join(HERE, "pcg_kats.json"), "w"), indent=1)
    json.dump(dict(material=64, sphere=80, cuboid=96, cuboid_base=20480, game_objects_ubo=26624, basic_data_ubo=144,
                   atmosphere_ubo=464, default_spheres=48, default_cuboids=7), open(os.path.join(HERE, "layout_pins.json"), "w"), indent=1)

    import ptb200
    from oracle import oracle as O
    sc = ptb200.scene
    env = O.atmosphere(16, sc.atmosphere_ubo_bytes(), sc.atmosphere_light_pos(0.5), 15.0, 8, 4)
    np.save(os.path.join(HERE, "env16.npy"), env)
    scene, cam = sc.load_default_scene(), sc.default_camera()
    basic, ubo = sc.basic_data_bytes(cam, 64, 64), scene.ubo_bytes()
    img = np.zeros((64, 64, 4), np.float32)
    for f in range(4):
        O.render(img, basic, ubo, env, frame=f, spp=1, ray_depth=13, focal_length=20.0, aperture_diameter=0.14, n_spheres=48, n_cuboids=7)
        if f == 0:
            np.save(os.path.join(HERE, "c1_64x64_f0.npy"), img)
    np.save(os.path.join(HERE, "c1_64x64_f0_3.npy"), img)
    np.save(os.path.join(HERE, "default_scene_ubo.npy"), np.frombuffer(ubo, dtype=np.uint8)[:48 * 80].copy())
    np.save(os.path.join(HERE, "default_scene_cuboids.npy"), np.frombuffer(ubo, dtype=np.uint8)[20480:20480 + 7 * 96].copy())
    np.save(os.path.join(HERE, "default_basic_ubo_64x64.npy"), np.frombuffer(basic, dtype=np.uint8).copy())
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
