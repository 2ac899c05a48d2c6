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


def _camera_from_vector(gb, v):
    return gb.Camera(imgsz=tuple(int(x) for x in v[6:8]), f=tuple(v[8:10]), c=tuple(v[10:12]), k=tuple(v[12:18]), p=tuple(v[18:20]),
                     xyz=tuple(v[0:3]), viewdir=tuple(v[3:6]))


@pytest.mark.parametrize("name", list(scenes.project_image_cases()))
def test_project_image_matches_reference(cuda, name):
    """Image.project (image.py:301-361) through gb_project_image against the reference's own output.  The device's rays differ
    from NumPy's by ~1e-12 px, so: the same pixels are seen / unseen; floating bands agree to 1e-9 of the value range (nearest:
    exactly, but for a sample that sits within 1e-9 px of a cell edge); integer bands are equal except where the interpolated
    value sits within 1e-6 of an integer (truncation), and then differ by one level."""
    import glimpse_b200 as gb
    from glimpse_b200 import synthetic

    dtype, bands, method, src, dst = scenes.project_image_cases()[name]
    ref = helpers.load_golden("project_image")[name]
    img = gb.Image(name, cam=_camera_from_vector(gb, src), datetime=synthetic.T0)
    img.array = scenes.project_image_frame(name)
    out = img.project(_camera_from_vector(gb, dst), method=method)
    assert out.dtype == ref.dtype and out.shape == ref.shape
    if np.issubdtype(dtype, np.integer):
        diff = np.abs(out.astype(np.int64) - ref.astype(np.int64))
        assert diff.max() <= (1 if method == "linear" else 0) or (method == "nearest" and (diff != 0).mean() < 1e-4)
        assert (diff != 0).mean() < 1e-4
    else:
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        ok = ~np.isnan(ref)
        span = float(np.nanmax(ref) - np.nanmin(ref))
        diff = np.abs(out[ok].astype(float) - ref[ok].astype(float))
        if method == "linear":
            assert diff.max() <= max(1e-9 * span, 2 * float(np.finfo(dtype).eps) * span)
        else:
            assert (diff != 0).mean() < 1e-4
    # and against the oracle on the same inputs (the checker the other tests use)
    mine = orc.project_image(scenes.project_image_frame(name), src, dst, method)
    np.testing.assert_array_equal(mine, ref)


def test_project_image_errors(cuda):
    import glimpse_b200 as gb
    from glimpse_b200 import synthetic

    _, _, _, src, dst = scenes.project_image_cases()["u8_rgb_linear"]
    img = gb.Image("x", cam=_camera_from_vector(gb, src), datetime=synthetic.T0)
    img.array = scenes.project_image_frame("u8_rgb_linear")
    moved = dst.copy()
    moved[2] += 0.5
    with pytest.raises(ValueError, match="different positions"):
        img.project(_camera_from_vector(gb, moved))
    with pytest.raises(ValueError, match="not defined"):
        img.project(_camera_from_vector(gb, dst), method="cubic")


@pytest.mark.parametrize("name", list(scenes.viewshed_cases()))
def test_viewshed_matches_reference(cuda, name):
    """Raster.viewshed (raster.py:1293-1389) through gb_viewshed against the reference's boolean array: equal cell for cell
    (the device's atan2 may differ from NumPy's in the last bit, which could only flip a cell whose elevation ratio sits within
    rounding of the interpolated horizon — none does in these cases)."""
    import warnings

    import glimpse_b200 as gb

    case = scenes.viewshed_cases()[name]
    ref = np.unpackbits(helpers.load_golden("viewshed")[name])[: case["z"].size].reshape(case["z"].shape).astype(bool)
    raster = gb.Raster(case["z"], x=case["xlim"], y=case["ylim"])
    corr = case["correction"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vis = raster.viewshed(case["origin"], correction=dict(radius=corr[0], refraction=corr[1]) if corr else False)
    assert vis.dtype == bool and vis.shape == ref.shape
    assert int((vis != ref).sum()) == 0, f"{int((vis != ref).sum())} of {ref.size} cells differ"


def test_viewshed_full_size_matches_reference(cuda):
    """2000 x 2000 cells, 1 400 rings of up to 8 900 cells: every one of the 4 000 000 cells as the reference has it."""
    import glimpse_b200 as gb

    case = scenes.viewshed_large_case()
    ref = np.unpackbits(helpers.load_golden("viewshed")["large_2000"])[: case["z"].size].reshape(case["z"].shape).astype(bool)
    raster = gb.Raster(case["z"], x=case["xlim"], y=case["ylim"])
    vis = raster.viewshed(case["origin"], correction=dict(radius=case["correction"][0], refraction=case["correction"][1]))
    assert int((vis != ref).sum()) == 0 and 0.005 < vis.mean() < 0.5


def test_viewshed_warnings_and_limits(cuda):
    import glimpse_b200 as gb

    z = scenes.viewshed_cases()["interior"]["z"]
    with pytest.warns(UserWarning, match="not square"):
        gb.Raster(z, x=(0.0, 800.0), y=(300.0, 0.0)).viewshed((400.0, 150.0, 5000.0))
    with pytest.warns(UserWarning, match="not in DEM"):
        gb.Raster(z, x=(0.0, 800.0), y=(600.0, 0.0)).viewshed((-10.0, 150.0, 5000.0))
    # from far above everything is visible; with the curvature correction too
    assert gb.Raster(z, x=(0.0, 800.0), y=(600.0, 0.0)).viewshed((400.0, 300.0, 1e6), correction=True).all()


def test_observer_sample_tile_and_shift_tile_match_fitpack(cuda):
    """Observer.sample_tile / shift_tile (reference observer.py:146-214) through gb_sample_surface: the device's Hermite-form
    not-a-knot spline against scipy's RectBivariateSpline (FITPACK) on a smooth tile — cubic, linear and mixed degrees, points
    and grids, arguments up to half a cell outside the outermost centres (evaluated at the clamped coordinate)."""
    import datetime

    import scipy.interpolate

    import glimpse_b200 as gb

    rng = np.random.RandomState(5)
    ny, nx = 23, 31
    yy, xx = np.mgrid[0:ny, 0:nx]
    tile = np.sin(xx / 4.0) * np.cos(yy / 5.0) + 0.05 * rng.rand(ny, nx)
    box = (100.0, 50.0, 100.0 + nx, 50.0 + ny)
    day = datetime.timedelta(days=1)
    cam = gb.Camera(imgsz=(200, 100), f=100.0)
    obs = gb.Observer([gb.Image(f"i{k}", cam=cam, datetime=datetime.datetime(2020, 1, 1) + k * day) for k in range(2)])
    uv = np.column_stack((box[0] + rng.rand(500) * nx, box[1] + rng.rand(500) * ny))
    cu, cv = np.arange(box[0] + 0.5, box[2]), np.arange(box[1] + 0.5, box[3])
    for kw in ({}, {"kx": 1, "ky": 1}, {"kx": 3, "ky": 1}, {"kx": 2, "ky": 2}, {"kx": 4, "ky": 5}, {"kx": 5, "ky": 2}):
        f = scipy.interpolate.RectBivariateSpline(cv, cu, tile, **kw)
        got = obs.sample_tile(uv, tile, box, **kw)
        assert np.max(np.abs(got - f(uv[:, 1], uv[:, 0], grid=False))) < 2e-6  # the device holds the tile as float32
        gu, gv = np.linspace(box[0], box[2], 17), np.linspace(box[1], box[3], 13)
        grid = obs.sample_tile((gu, gv), tile, box, grid=True, **kw)
        assert grid.shape == (13, 17) and np.max(np.abs(grid - f(gv, gu, grid=True))) < 2e-6
    with pytest.raises(ValueError, match="outside box"):
        obs.sample_tile(np.array([[box[0] - 0.1, box[1] + 1.0]]), tile, box)
    duv = (0.3, -0.45)
    f = scipy.interpolate.RectBivariateSpline(np.arange(0.5, ny), np.arange(0.5, nx), tile)
    want = f(np.arange(0.5, ny) + duv[1], np.arange(0.5, nx) + duv[0], grid=True)
    assert np.max(np.abs(obs.shift_tile(tile.copy(), duv) - want)) < 2e-6
    rgb = np.dstack([tile, 2 * tile, -tile])
    assert np.max(np.abs(obs.shift_tile(rgb, duv)[:, :, 1] - 2 * want)) < 4e-6


def _jpeg_bytes(array, quality=92, full_chroma=False):
    import cv2

    params = [int(cv2.IMWRITE_JPEG_QUALITY), quality]
    if full_chroma:  # 4:4:4: no chroma subsampling, so that decoders only differ in their IDCT and colour conversion
        params += [int(cv2.IMWRITE_JPEG_SAMPLING_FACTOR), int(cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444)]
    ok, buf = cv2.imencode(".jpg", array[:, :, ::-1] if array.ndim == 3 else array, params)
    assert ok
    return buf.tobytes()


def _decode_or_skip(data, **kw):
    from glimpse_b200.image import decode_jpeg

    try:
        return decode_jpeg(data, **kw)
    except Exception as exc:  # the library is opened at run time
        if "libnvjpeg" in str(exc):
            pytest.skip("nvJPEG is not installed on this box")
        raise


@pytest.mark.parametrize("bands", [3, 1])
def test_jpeg_decode_on_the_device_is_close_to_libjpeg(cuda, bands):
    """Frame ingest (image.py:137-214): JPEG bytes decoded by nvJPEG straight into device memory.  Decoders differ in IDCT and
    chroma upsampling, so the check is closeness to libjpeg-turbo's decode of the same stream (OpenCV), not identity."""
    import cv2
    from glimpse_b200 import synthetic

    rng = np.random.RandomState(5)
    tex = synthetic.smooth_texture((240, 320), rng)
    frame = np.stack([tex, np.roll(tex, 5, axis=1), np.roll(tex, -3, axis=0)], axis=2) if bands == 3 else tex
    data = _jpeg_bytes(np.ascontiguousarray(frame), full_chroma=True)
    got = _decode_or_skip(data).cpu().numpy()
    ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_UNCHANGED)
    ref = ref[:, :, ::-1] if ref.ndim == 3 else ref
    assert got.shape == ref.shape == frame.shape and got.dtype == np.uint8
    diff = np.abs(got.astype(int) - ref.astype(int))
    assert diff.max() <= 4 and diff.mean() < 0.75, (diff.max(), diff.mean())
    assert np.abs(got.astype(int) - frame.astype(int)).mean() < 3.0  # and both are the picture that was encoded
    if bands == 3:
        # a subsampled (4:2:0) stream of a picture whose colour changes from pixel to pixel: the decoders interpolate the
        # chroma planes differently (libjpeg's triangle filter, nvJPEG's replication), the luma they agree on
        sub = _jpeg_bytes(np.ascontiguousarray(frame))
        y = _decode_or_skip(sub, gray=True).cpu().numpy()
        y_ref = cv2.cvtColor(cv2.imdecode(np.frombuffer(sub, np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2YCrCb)[:, :, 0]
        assert y.shape == frame.shape[:2]
        d = np.abs(y.astype(int) - y_ref.astype(int))
        assert d.max() <= 4 and d.mean() < 0.75, (d.max(), d.mean())


def test_tracking_from_device_decoded_frames(cuda, tmp_path):
    """Observer.cache_images(device=True): the JPEG files are decoded on the device and tracked from there (no upload of pixel
    arrays) — the same track, bit for bit, as from host arrays holding those very pixels."""
    import glimpse_b200 as gb
    from glimpse_b200 import synthetic

    scene = synthetic.nadir_scene(seed=8, n_points=4, n_particles=600, n_frames=6, imgsz=(320, 240), margin_px=90, bands=3)
    observers, models = synthetic.build(scene, gb)
    images = observers[0].images
    for i, img in enumerate(images):
        path = tmp_path / f"frame{i}.jpg"
        path.write_bytes(_jpeg_bytes(np.ascontiguousarray(img.array), quality=95))
        img.path, img.array = str(path), None
    try:
        observers[0].cache_images(device=True)
    except Exception as exc:
        if "libnvjpeg" in str(exc):
            pytest.skip("nvJPEG is not installed on this box")
        raise
    assert all(img.device_array is not None and img.array is None for img in images)
    on_device = gb.Tracker(observers, seed=3)
    a = on_device.track(models, tile_size=scene.tile_size)
    assert all(e is None for e in a.errors)
    # the same pixels as host arrays
    for img in images:
        img.array = img.device_array.cpu().numpy()
        img.device_array = None
    from_host = gb.Tracker(observers, seed=3)
    b = from_host.track(models, tile_size=scene.tile_size)
    np.testing.assert_array_equal(a.means, b.means)
    np.testing.assert_array_equal(a.sigmas, b.sigmas)
    assert on_device.last_run["h2d_bytes"] < from_host.last_run["h2d_bytes"] - 5 * 320 * 240 * 3
    # and the decoded sequence still carries the motion
    v = a.vxyz[:, -1, 0]
    assert np.all(np.abs(v - scene.truth_velocity[0]) < 0.1)
