"""Scene and camera definitions shared by the golden generator and the tests.

Everything is seeded and rebuilt identically on both sides, so the fixtures in ``tests/golden``
only need to carry the reference's outputs.
"""
import datetime

import numpy as np

from glimpse_b200 import synthetic


def camera_configs():
    """Distortion cases of the reference's ``tests/test_camera.py:34-88`` on a map-scale camera,
    plus a curvature/refraction case (``camera.py:1444-1449``)."""
    base = dict(imgsz=(4288, 2848), f=(3700.0, 3690.0), c=(12.5, -8.25), xyz=(4.99e5, 6.77e6, 500.0),
                viewdir=(60.0, -25.0, 1.5))
    small = dict(imgsz=(100, 100), f=100.0, xyz=(1.0, 2.0, 3.0), viewdir=(10.0, 20.0, 30.0))
    return {
        "ideal": dict(base),
        "k1": dict(base, k=(0.1, 0, 0, 0, 0, 0)),
        "k1neg": dict(base, k=(-0.1, 0, 0, 0, 0, 0)),
        "k123": dict(base, k=(0.05, -0.01, 0.001, 0, 0, 0)),
        "k456": dict(base, k=(0, 0, 0, 0.002, 0.0005, -0.0001)),
        "k6": dict(base, k=(0.1,) * 6),
        "p": dict(base, p=(0.01, 0.01)),
        "full": dict(base, k=synthetic.FULL_K, p=synthetic.FULL_P),
        "full_corr": dict(base, k=synthetic.FULL_K, p=synthetic.FULL_P, correction=True),
        "small_k1_extreme": dict(small, k=(2.0, 0, 0, 0, 0, 0)),
        "small_k1_neg_extreme": dict(small, k=(-2.0, 0, 0, 0, 0, 0)),
        "small_all": dict(small, k=(0.1,) * 6, p=(0.01, 0.01)),
    }


def viewshed_cases():
    """``Raster.viewshed`` (raster.py:1293-1389) cases: name -> dict(z (ny, nx), xlim, ylim, origin, correction = (radius, refraction)
    or None).  Bumpy seeded DEMs; the goldens (tests/golden/viewshed.npz) hold the reference's boolean arrays."""
    from scipy.ndimage import gaussian_filter

    def dem(seed, ny, nx, relief=200.0):
        rng = np.random.RandomState(seed)
        return gaussian_filter(rng.rand(ny, nx), 3) * relief * 8 + np.linspace(0, 30, nx)[None, :]

    def above(z, xlim, ylim, x, y, h):  # an eye h above the cell under (x, y)
        col = int((x - xlim[0]) / ((xlim[1] - xlim[0]) / z.shape[1]))
        row = int((y - ylim[0]) / ((ylim[1] - ylim[0]) / z.shape[0]))
        return (x, y, float(np.nanmax(z[max(row - 1, 0):row + 2, max(col - 1, 0):col + 2])) + h)

    xl, yl = (100.0, 900.0), (700.0, 100.0)
    z = dem(3, 60, 80)
    holes = z.copy()
    holes[10:14, 20:30] = np.nan
    holes[40, 5] = np.nan
    holes[29:32, 38:41] = np.nan  # around the eye: the first ring has NaN cells
    big = dem(5, 300, 420, relief=120.0)
    return {
        "interior": dict(z=z, xlim=xl, ylim=yl, origin=above(z, xl, yl, 483.3, 391.2, 3.0), correction=None),
        "corner": dict(z=z, xlim=xl, ylim=yl, origin=above(z, xl, yl, 105.0, 695.0, 2.0), correction=None),
        "on_cell_corner_corrected": dict(z=z, xlim=xl, ylim=yl, origin=above(z, xl, yl, 500.0, 400.0, 5.0), correction=(6.3781e6, 0.13)),
        "on_cell_centre": dict(z=z, xlim=xl, ylim=yl, origin=above(z, xl, yl, 505.0, 395.0, 1.5), correction=None),  # a ring 0 of one cell
        "nan_cells": dict(z=holes, xlim=xl, ylim=yl, origin=above(z, xl, yl, 483.3, 391.2, 3.0), correction=None),
        "single_cell": dict(z=z[:1, :1], xlim=(0.0, 10.0), ylim=(10.0, 0.0), origin=(5.0, 5.0, 1000.0), correction=None),
        "outside_flipped": dict(z=z[:30, :40], xlim=(900.0, 100.0), ylim=(100.0, 700.0), origin=(-50.0, 50.0, float(z.max()) + 40.0), correction=None),
        "one_row": dict(z=z[:1, :], xlim=xl, ylim=(110.0, 100.0), origin=(300.0, 105.0, float(z[0].max()) - 5.0), correction=None),
        "big": dict(z=big, xlim=(0.0, 4200.0), ylim=(3000.0, 0.0), origin=above(big, (0.0, 4200.0), (3000.0, 0.0), 1711.0, 1203.0, 150.0),
                    correction=(6.3781e6, 0.13)),
    }


def viewshed_large_case(side=2000):
    """A ``side`` x ``side`` DEM (bilinear blow-up of seeded noise, 900 m of relief, 10 m cells) seen from 30 m above a point
    near its middle, curvature / refraction corrected: the full-size parity and timing case (golden ``viewshed_2000``)."""
    rng = np.random.RandomState(7)
    coarse = rng.rand(side // 50 + 2, side // 50 + 2)
    yy, xx = np.mgrid[0:side, 0:side] / 50.0
    i, j = yy.astype(int), xx.astype(int)
    fy, fx = yy - i, xx - j
    z = ((coarse[i, j] * (1 - fy) + coarse[i + 1, j] * fy) * (1 - fx) + (coarse[i, j + 1] * (1 - fy) + coarse[i + 1, j + 1] * fy) * fx) * 900.0
    mid = side // 2
    origin = (side * 5.0 + 3.0, side * 5.0 - 4.0, float(z[mid - 2:mid + 3, mid - 2:mid + 3].max()) + 30.0)
    return dict(z=z, xlim=(0.0, side * 10.0), ylim=(side * 10.0, 0.0), origin=origin, correction=(6.3781e6, 0.13))


def project_image_cases():
    """``Image.project`` (image.py:301-361) cases: name -> (frame dtype, bands, method, source camera vector, target camera vector).
    The frames are ``project_image_frame(name)``; the goldens (tests/golden/project_image.npz) hold the reference's output."""
    v = synthetic.camera_vector
    pos = (10.0, 20.0, 30.0)
    src = v(imgsz=(160, 120), f=(300.0, 295.0), xyz=pos, viewdir=(30.0, -10.0, 2.0), c=(1.5, -2.0), k=synthetic.FULL_K, p=synthetic.FULL_P)
    dst = v(imgsz=(140, 100), f=(280.0, 280.0), xyz=pos, viewdir=(32.0, -9.0, 0.0), k=(0.1, 0, 0, 0, 0, 0))
    same = v(imgsz=(160, 120), f=(300.0, 295.0), xyz=pos, viewdir=(31.0, -10.5, 1.0))  # same frame size: the grids are shared
    wide = v(imgsz=(200, 90), f=(150.0, 150.0), xyz=pos, viewdir=(30.0, -10.0, 2.0))   # sees far beyond the source: NaN border
    return {
        "u8_rgb_linear": (np.uint8, 3, "linear", src, dst),
        "u8_gray_nearest": (np.uint8, 1, "nearest", src, dst),
        "u8_gray_linear_same": (np.uint8, 1, "linear", src, same),
        "u16_rgb_linear": (np.uint16, 3, "linear", src, dst),
        "f32_gray_linear": (np.float32, 1, "linear", src, wide),
        "f32_rgb_nearest": (np.float32, 3, "nearest", src, same),
        "f64_rgb_linear": (np.float64, 3, "linear", src, dst),
        "f64_gray_nearest_wide": (np.float64, 1, "nearest", src, wide),
        "i16_gray_linear": (np.int16, 1, "linear", src, dst),  # a type the device holds promoted to float64
    }


def project_image_frame(name):
    dtype, bands, _, src, _ = project_image_cases()[name]
    rng = np.random.RandomState(sum(map(ord, name)))
    w, h = (int(x) for x in src[6:8])
    tex = rng.rand(h, w, bands) * 250
    if dtype == np.int16:
        tex = tex * 40 - 3000
    elif dtype == np.uint16:
        tex = tex * 200
    elif np.issubdtype(dtype, np.floating):
        tex = tex / 250
    frame = tex.astype(dtype)
    return frame if bands > 1 else frame[:, :, 0]


def add_second_observer(scene):
    """Second station: same place, rolled 180 deg (image flipped both ways), RGB frames, radial-only
    distortion, and images starting one frame later (staggered template start)."""
    first = scene.observers[0]
    frames, cams, dts = [], [], []
    for t in range(1, len(first.frames)):
        g = first.frames[t][::-1, ::-1]
        rgb = np.stack([g, np.roll(g, 1, axis=1), np.roll(g, 1, axis=0)], axis=2)
        frames.append(np.ascontiguousarray(rgb))
        vec = first.cams[t].copy()
        vec[3:6] = (0.0, -90.0, 180.0)
        vec[18:20] = 0.0
        cams.append(vec)
        dts.append(first.datetimes[t])
    scene.observers.append(synthetic.ObserverScene(frames, np.array(cams), dts, sigma=0.4))
    return scene


def add_gridded_dem(scene):
    """A gently sloping, bumpy DEM and a spatially varying DEM sigma over the whole footprint (north-up raster:
    y decreasing), as dict(array, x, y) that ``synthetic.build`` turns into the API's Raster."""
    rng = np.random.RandomState(11)
    ny, nx = 14, 18
    xs = np.linspace(-1, 1, nx)[None, :]
    ys = np.linspace(-1, 1, ny)[:, None]
    dem = 0.8 * xs + 0.4 * ys + 0.15 * rng.rand(ny, nx)
    sig = 0.2 + 0.1 * rng.rand(ny, nx)
    x, y = (-45.0, 45.0), (35.0, -35.0)
    scene.motion.update(dem=dict(array=dem, x=x, y=y), dem_sigma=dict(array=sig, x=x, y=y))
    return scene


def narrow_cloud(scene):
    """Particles that stay within a fraction of a pixel: the search window is widened to the minimum the spline degree
    needs (tracker.py:584-594), so the SSE surface has only degree + 1 cells per axis."""
    scene.motion.update(xy_sigma=(0.01, 0.01), vxyz_sigma=(0.01, 0.01, 0.0), axyz_sigma=(0.002, 0.002, 0.0))
    return scene


def frames_as(dtype):
    """Post-processing of a scene: the same frames in another pixel type, with sub-grey-level detail so that a window holds far
    more distinct values than the 256 of the uint8 frames (uint16: 200 sub-levels; float32 / float64: grey / 255 plus noise)."""
    def convert(scene):
        rng = np.random.RandomState(17)
        for obs in scene.observers:
            frames = []
            for f in obs.frames:
                if dtype == np.uint16:
                    g = f.astype(np.uint16) * 200 + rng.randint(0, 200, f.shape).astype(np.uint16)
                else:
                    g = ((f.astype(np.float64) + rng.rand(*f.shape)) / 255.0).astype(dtype)
                frames.append(np.ascontiguousarray(g))
            obs.frames = frames
        return scene
    return convert


def chain(*steps):
    def run(scene):
        for step in steps:
            scene = step(scene)
        return scene
    return run


def track_cases():
    return {
        # config-1 shape, shrunk: 1 observer, Cartesian, full distortion
        "track_c1": dict(
            scene_kwargs=dict(seed=1, n_points=3, n_particles=300, n_frames=6, imgsz=(320, 240), margin_px=90),
            seed=101,
        ),
        # config-3 shape, shrunk: 2 observers (RGB second, staggered), Cylindrical, dem_sigma > 0, covariances
        "track_cyl2": dict(
            scene_kwargs=dict(seed=3, n_points=2, n_particles=256, n_frames=5, imgsz=(320, 240), margin_px=100,
                              kind="cylindrical", velocity_sigma=0.2),
            seed=303, post=add_second_observer, return_covariances=True,
        ),
        # tangent models on a gridded, sloping DEM (two bilinear DEM gathers per particle and step)
        "track_tangent": dict(
            scene_kwargs=dict(seed=7, n_points=2, n_particles=256, n_frames=5, imgsz=(320, 240), margin_px=100,
                              kind="tangent_cartesian"),
            seed=707, post=add_gridded_dem,
        ),
        "track_tangent_cyl": dict(
            scene_kwargs=dict(seed=9, n_points=2, n_particles=256, n_frames=5, imgsz=(320, 240), margin_px=100,
                              kind="tangent_cylindrical", velocity_sigma=0.2),
            seed=909, post=add_gridded_dem,
        ),
        # stratified resampling, one uniform per particle and update (tracker.py:178-186)
        "track_stratified": dict(
            scene_kwargs=dict(seed=13, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=1313, resample_method="stratified",
        ),
        "track_choice": dict(
            scene_kwargs=dict(seed=15, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=1515, resample_method="choice",
        ),
        # frames other than uint8 (tracker.py:522-524 takes any dtype): the rank pipeline of the device
        "track_u16": dict(
            scene_kwargs=dict(seed=47, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=4747, post=frames_as(np.uint16),
        ),
        "track_u16_rgb": dict(  # second observer with three uint16 bands, covariances
            scene_kwargs=dict(seed=49, n_points=2, n_particles=256, n_frames=5, imgsz=(320, 240), margin_px=100,
                              kind="cylindrical", velocity_sigma=0.2),
            seed=4949, post=chain(frames_as(np.uint16), add_second_observer), return_covariances=True,
        ),
        "track_f32": dict(
            scene_kwargs=dict(seed=51, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=5151, post=frames_as(np.float32),
        ),
        "track_f64_hp3": dict(  # float64 frames, 3 x 3 median on a 'nearest' border
            scene_kwargs=dict(seed=53, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=5353, post=frames_as(np.float64), highpass={"size": 3, "mode": "nearest"},
        ),
        # Raster observers: an Observer whose frames are orthoimages (Raster with a datetime; xyz_to_uv is the grid's affine map,
        # raster.py:423-445) — alone, and as a float32 orthoimage next to an RGB camera station
        "track_raster": dict(
            scene_kwargs=dict(seed=55, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=5555, post=synthetic.as_raster_frames,
        ),
        "track_raster_mixed": dict(
            scene_kwargs=dict(seed=57, n_points=2, n_particles=256, n_frames=5, imgsz=(320, 240), margin_px=100,
                              kind="cylindrical", velocity_sigma=0.2),
            seed=5757, post=chain(frames_as(np.float32), add_second_observer, synthetic.as_raster_frames), return_covariances=True,
        ),
        # residual resampling as the reference computes it (tracker.py:188-203); wide velocity prior so that weights are uneven
        "track_residual": dict(
            scene_kwargs=dict(seed=37, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=3737, resample_method="residual",
        ),
        # other sizes of the median high-pass (Tracker.highpass, tracker.py:59, 530):
        # rows != columns, and an even size given as one integer (window offsets -2 .. 1)
        "track_hp37": dict(
            scene_kwargs=dict(seed=17, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=1717, highpass={"size": (3, 7)},
        ),
        "track_hp4": dict(
            scene_kwargs=dict(seed=19, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=1919, post=add_second_observer, highpass={"size": 4},
        ),
        # ... other border modes and origins of the median filter (scipy.ndimage.median_filter's mode / cval / origin)
        "track_hp_mirror": dict(
            scene_kwargs=dict(seed=29, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=2929, highpass={"size": 5, "mode": "mirror"},
        ),
        "track_hp_const": dict(
            scene_kwargs=dict(seed=31, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=3131, post=add_second_observer, highpass={"size": (3, 5), "mode": "constant", "cval": 0.1, "origin": (1, -2)},
        ),
        "track_hp_wrap": dict(
            scene_kwargs=dict(seed=33, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=3333, highpass={"size": (4, 6), "mode": "wrap", "origin": (-2, 2)},
        ),
        "track_hp_nearest": dict(
            scene_kwargs=dict(seed=35, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=3535, highpass={"size": 7, "mode": "nearest", "origin": 3},
        ),
        "track_hp_footprint": dict(  # a plus-shaped footprint with a hole, shifted, on a wrapped border
            scene_kwargs=dict(seed=39, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=3939, highpass={"footprint": [[0, 0, 1, 0, 0], [0, 1, 1, 1, 0], [1, 1, 0, 1, 1], [0, 1, 1, 1, 0]], "mode": "mirror",
                                 "origin": (0, 1)},
        ),
        # ... and other spline degrees (Tracker.interpolation, tracker.py:60): bilinear; a 2 x 2 .. 3 x 3 surface (bilinear,
        # window widened to the minimum); cubic along the rows with linear along the columns, two observers
        "track_lin": dict(
            scene_kwargs=dict(seed=21, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=2121, interpolation={"kx": 1, "ky": 1},
        ),
        "track_lin_narrow": dict(
            scene_kwargs=dict(seed=23, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=2323, post=narrow_cloud, interpolation={"kx": 1, "ky": 1},
        ),
        "track_k31": dict(
            scene_kwargs=dict(seed=25, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=2525, post=add_second_observer, interpolation={"kx": 3, "ky": 1},
        ),
        # degrees 2, 4, 5 (FITPACK's knots between the data sites for even degrees): B-spline coefficients on the device
        "track_k22": dict(
            scene_kwargs=dict(seed=41, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=4141, interpolation={"kx": 2, "ky": 2},
        ),
        "track_k45": dict(
            scene_kwargs=dict(seed=43, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=4343, post=add_second_observer, interpolation={"kx": 4, "ky": 5},
        ),
        "track_k52_narrow": dict(  # surfaces of 6 x 3 cells
            scene_kwargs=dict(seed=45, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=4545, post=narrow_cloud, interpolation={"kx": 5, "ky": 2},
        ),
        "track_narrow": dict(  # the default cubic on a widened 4 x 4 surface
            scene_kwargs=dict(seed=27, n_points=2, n_particles=300, n_frames=5, imgsz=(320, 240), margin_px=100),
            seed=2727, post=narrow_cloud,
        ),
        # map-scale world coordinates + per-frame view-direction jitter
        "track_jitter": dict(
            scene_kwargs=dict(seed=5, n_points=2, n_particles=256, n_frames=5, imgsz=(320, 240), margin_px=100,
                              jitter_deg=0.02, world_offset=(4.99e5, 6.77e6)),
            seed=505,
        ),
    }


def shape_cases():
    """Parity fixtures at the shapes BASELINE.json names (full particle counts, tile sizes, frame sizes and distortion; fewer
    points / frames for configs 2-4 so that the CPU reference finishes in seconds).  Their goldens are compact: means, sigmas,
    every update's ancestor indices, uniform draw and checksums of the evolved particles / weights, one SSE surface."""
    return {
        # config 1 exactly: 10 points x 1 000 particles x 20 frames, 15 x 15 template, 600 x 400 frames
        "shape_c1": dict(
            scene_kwargs=dict(seed=41, n_points=10, n_particles=1000, n_frames=20, imgsz=(600, 400)),
            seed=4141,
        ),
        # config 2: N = 10 000, 4288 x 2848 nadir camera with the full k1-k6 / p1 / p2 distortion
        "shape_c2": dict(
            scene_kwargs=dict(seed=42, n_points=3, n_particles=10000, n_frames=8, imgsz=(4288, 2848), velocity_sigma=0.2,
                              margin_px=200),
            seed=4242,
        ),
        # config 3: CylindricalMotion with an uncertain elevation (dem_sigma = 1), two observers (the second rolled by 180 deg,
        # RGB, radial-only distortion, starting one frame later), N = 10 000
        "shape_c3": dict(
            scene_kwargs=dict(seed=43, n_points=2, n_particles=10000, n_frames=6, imgsz=(1200, 800), kind="cylindrical",
                              velocity_sigma=0.2, margin_px=200),
            seed=4343, post=add_second_observer,
        ),
        # config 4: 31 x 31 template, N = 100 000 (search windows of ~100 px)
        "shape_c4": dict(
            scene_kwargs=dict(seed=44, n_points=2, n_particles=100000, n_frames=5, imgsz=(1200, 800), tile_size=(31, 31),
                              velocity_sigma=0.3, margin_px=250),
            seed=4444,
        ),
    }
