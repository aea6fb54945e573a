import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


def f32_same(a, b):
    """Bitwise float32 equality, with any NaN equal to any NaN (x86 and sm_100 differ in NaN payload)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ptb():
    import ptb200
    return ptb200


@pytest.fixture(scope="session")
def env256(oracle, ptb):
    """The default environment: AtmosphericScatterer(256), Time 0.5, 50x15 steps, intensity 15 (MainWindow.cs:174, AtmosphericScatterer.cs:91-94)."""
    sc = ptb.scene
    return oracle.atmosphere(256, sc.atmosphere_ubo_bytes(), sc.atmosphere_light_pos(0.5), 15.0, 50, 15)


@pytest.fixture(scope="session")
def env16(oracle, ptb):
    sc = ptb.scene
    return oracle.atmosphere(16, sc.atmosphere_ubo_bytes(), sc.atmosphere_light_pos(0.5), 15.0, 8, 4)


@pytest.fixture(scope="session")
def default_scene(ptb):
    return ptb.load_default_scene()


@pytest.fixture(scope="session")
def camera(ptb):
    return ptb.default_camera()


def oracle_render(O, sc_mod, scene, camera, env, W, H, frames, *, spp=1, depth=13, focal=20.0, aperture=0.14, rows=None,
                  cols=None, first_frame=0, image=None):
    basic = sc_mod.basic_data_bytes(camera, W, H)
    ubo = scene.ubo_bytes()
    img = np.zeros((H, W, 4), np.float32) if image is None else image
    for f in range(first_frame, first_frame + frames):
        O.render(img, basic, ubo, env, frame=f, spp=spp, ray_depth=depth, focal_length=focal, aperture_diameter=aperture,
                 n_spheres=len(scene.spheres), n_cuboids=len(scene.cuboids), max_spheres=scene.max_spheres, rows=rows, cols=cols)
    return img
