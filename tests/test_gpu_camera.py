"""Projection / inverse projection kernels vs vectors produced by the reference Camera
(tests/golden/camera.npz, made by tests/golden/make_golden.py) and vs the oracle."""
import numpy as np
import pytest

import helpers
import scenes
from oracle import tracker_oracle as orc

pytestmark = pytest.mark.gpu

UV_TOL_PX = 1e-9      # fp64 projection; north_star's outer bound is 1e-4 px
DIR_TOL = 1e-11       # ray directions (unit depth)


def make_camera(kw):
    import glimpse_b200 as gb

    kw = dict(kw)
    corr = kw.pop("correction", False)
    return gb.Camera(correction=corr, **kw)


@pytest.mark.parametrize("name", list(scenes.camera_configs()))
def test_project_matches_reference(cuda, name):
    g = helpers.load_golden("camera")
    cam = make_camera(scenes.camera_configs()[name])
    np.testing.assert_array_equal(cam.vector, g[f"{name}.vector"])
    uv = cam.xyz_to_uv(g[f"{name}.xyz"])
    ref = g[f"{name}.uv"]
    assert np.array_equal(np.isnan(uv), np.isnan(ref))
    assert np.isnan(ref[-4:]).all()  # points behind the camera
    ok = ~np.isnan(ref)
    assert np.max(np.abs(uv[ok] - ref[ok])) <= UV_TOL_PX


@pytest.mark.parametrize("name", list(scenes.camera_configs()))
def test_unproject_matches_reference(cuda, name):
    g = helpers.load_golden("camera")
    cam = make_camera(scenes.camera_configs()[name])
    dirs = cam.uv_to_xyz(g[f"{name}.uv_in"])
    ref = g[f"{name}.dirs"]
    ok = ~np.isnan(ref).any(axis=1)
    assert np.array_equal(np.isnan(dirs).any(axis=1), ~ok)
    assert np.max(np.abs(dirs[ok] - ref[ok])) <= DIR_TOL


@pytest.mark.parametrize("name", ["ideal", "k1", "k6", "p", "full", "small_all", "small_k1_extreme"])
def test_reprojection_round_trip(cuda, name):
    """reference tests/test_camera.py:34-88: uv -> xyz -> uv returns to the start (ideal 1e-14 there;
    distorted 1e-12 in camera units; here in pixels on a 4288 px frame)."""
    cam = make_camera(scenes.camera_configs()[name])
    rng = np.random.RandomState(3)
    uv = rng.rand(1000, 2) * cam.imgsz
    xyz = cam.uv_to_xyz(uv, directions=False, depth=rng.rand(1000) * 1000 + 100)
    back = cam.xyz_to_uv(xyz)
    # map-scale coordinates (6.77e6 m) leave ~1e-9 m of fp64 cancellation in xyz - cam.xyz -> ~1e-8 px
    tol = 1e-7 if name in ("ideal",) else 2e-7
    assert np.nanmax(np.abs(back - uv)) < tol


def test_doctest_known_answers(cuda):
    """reference camera.py:615-620, 683-694."""
    import glimpse_b200 as gb

    cam = gb.Camera(imgsz=10, f=10)
    np.testing.assert_allclose(cam.xyz_to_uv(np.array([(0.0, 10.0, 0.0)])), [[5.0, 5.0]], atol=1e-14)
    np.testing.assert_allclose(cam.uv_to_xyz(np.array([(5.0, 5.0)])), [[0.0, 1.0, 0.0]], atol=1e-14)
    np.testing.assert_allclose(cam.uv_to_xyz(np.array([(5.0, 5.0)]), depth=10), [[0.0, 10.0, 0.0]], atol=1e-13)
    uv = cam.xyz_to_uv(np.array([(1000.0, 10, 0), (0, 10, 0), (0, 0, 0), (0, -10, 0)]))
    np.testing.assert_allclose(uv[:2], [[1005.0, 5.0], [5.0, 5.0]], atol=1e-12)
    assert np.isnan(uv[2:]).all()
    np.testing.assert_array_equal(cam.inframe(uv), [False, True, False, False])


def test_project_large_batch_against_oracle(cuda):
    cam = make_camera(scenes.camera_configs()["full_corr"])
    rng = np.random.RandomState(11)
    uv = rng.rand(200000, 2) * cam.imgsz
    xyz = orc.unproject(cam.vector, uv, directions=False, depth=200 + rng.rand(200000) * 5000)
    got = cam.xyz_to_uv(xyz)
    want = orc.project(cam.vector, xyz, correction=(cam.correction["radius"], cam.correction["refraction"]))
    assert np.max(np.abs(got - want)) <= UV_TOL_PX
