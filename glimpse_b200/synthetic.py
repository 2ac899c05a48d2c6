"""Seeded synthetic scenes for the Tracker hot path.

A scene is plain NumPy + datetimes: per observer a list of uint8 frames cut from one smooth
random texture that translates by a fixed number of pixels per frame, the per-frame camera
20-vectors, and a grid of tracked points with their motion-model parameters.  The same scene
can be turned into this package's objects (``build(scene, glimpse_b200)``) or — in the build
container, for golden vectors — into the reference's objects (``build(scene, glimpse)``),
because the constructors are signature-compatible (reference ``camera.py:74-123``,
``image.py:86-119``, ``track/observer.py:50-69``, ``track/motion.py:122-147,239-258``).
"""
from __future__ import annotations

import datetime
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

DAY = datetime.timedelta(days=1)
T0 = datetime.datetime(2020, 1, 1)

FULL_K = (0.05, -0.01, 0.001, 0.002, 0.0005, -0.0001)
FULL_P = (0.001, -0.0005)


def camera_vector(imgsz, f, xyz=(0, 0, 0), viewdir=(0, 0, 0), c=(0, 0), k=(0,) * 6, p=(0, 0)) -> np.ndarray:
    """The reference's 20-float camera vector [xyz, viewdir, imgsz, f, c, k1..k6, p1, p2]."""
    f = np.broadcast_to(np.asarray(f, float), (2,))
    return np.concatenate([np.asarray(xyz, float), np.asarray(viewdir, float), np.asarray(imgsz, float), f,
                           np.asarray(c, float), np.asarray(k, float), np.asarray(p, float)])


def smooth_texture(shape: Tuple[int, int], rng: np.random.RandomState, sigma: float = 1.5) -> np.ndarray:
    """uint8 texture = gaussian-filtered white noise stretched to 0..255."""
    import scipy.ndimage

    tex = scipy.ndimage.gaussian_filter(rng.rand(*shape), sigma=sigma)
    tex -= tex.min()
    tex *= 255.0 / tex.max()
    return tex.astype(np.uint8)


@dataclass
class ObserverScene:
    frames: List[np.ndarray]
    cams: np.ndarray  # (n_images, 20)
    datetimes: List[datetime.datetime]
    sigma: float = 0.3


@dataclass
class Scene:
    observers: List[ObserverScene]
    points: np.ndarray  # (P, 2) world xy
    motion: Dict  # kind + keyword parameters (without xy / n / time_unit)
    n_particles: int
    tile_size: Tuple[int, int] = (15, 15)
    time_unit: datetime.timedelta = DAY
    truth_velocity: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    meta: Dict = field(default_factory=dict)

    @property
    def datetimes(self) -> np.ndarray:
        return np.unique(np.concatenate([np.asarray(o.datetimes) for o in self.observers]))


def nadir_scene(
    seed: int,
    n_points: int,
    n_particles: int,
    n_frames: int,
    imgsz: Tuple[int, int] = (600, 400),
    shift_px: Tuple[int, int] = (2, 0),
    metres_per_px: float = 0.2,
    height: float = 1000.0,
    distortion: bool = True,
    velocity_sigma: float = 0.3,
    tile_size: Tuple[int, int] = (15, 15),
    margin_px: int = 120,
    bands: int = 1,
    kind: str = "cartesian",
    jitter_deg: float = 0.0,
    world_offset: Tuple[float, float] = (0.0, 0.0),
) -> Scene:
    """One nadir observer above a flat surface (z = 0) whose texture drifts ``shift_px`` per frame.

    Camera: ``xyz=(ox, oy, height)``, ``viewdir=(0, -90, 0)``, ``f = height / metres_per_px`` so one
    pixel is ``metres_per_px`` on the ground; with ``distortion`` the full k1-k6/p1/p2 model.  The
    true ground velocity is ``shift_px * metres_per_px`` per day along (+x, -y)."""
    rng = np.random.RandomState(seed)
    W, H = imgsz
    pad = max(abs(shift_px[0]), abs(shift_px[1])) * n_frames + 2
    tex = smooth_texture((H + 2 * pad, W + 2 * pad), rng)
    if bands == 3:
        tex = np.stack([tex, np.roll(tex, 3, axis=1), np.roll(tex, -2, axis=0)], axis=2)
    frames = []
    for t in range(n_frames):
        r0, c0 = pad - t * shift_px[1], pad - t * shift_px[0]
        frames.append(np.ascontiguousarray(tex[r0:r0 + H, c0:c0 + W]))
    f = height / metres_per_px
    cams = np.empty((n_frames, 20))
    for t in range(n_frames):
        jitter = rng.randn(3) * jitter_deg if jitter_deg else np.zeros(3)
        cams[t] = camera_vector(
            imgsz=(W, H), f=f, xyz=(world_offset[0], world_offset[1], height),
            viewdir=np.array((0.0, -90.0, 0.0)) + jitter,
            k=FULL_K if distortion else (0,) * 6, p=FULL_P if distortion else (0, 0),
        )
    # Regular grid of points inside the footprint, margin_px away from the frame edge
    nx = int(np.ceil(np.sqrt(n_points * W / H)))
    ny = int(np.ceil(n_points / nx))
    us = np.linspace(margin_px, W - margin_px, nx) if nx > 1 else np.array([W / 2.0])
    vs = np.linspace(margin_px, H - margin_px, ny) if ny > 1 else np.array([H / 2.0])
    uu, vv = np.meshgrid(us, vs)
    uv = np.column_stack((uu.ravel(), vv.ravel()))[:n_points]
    # Invert the (undistorted) nadir projection: u = f x / h + W/2, v = -f y / h + H/2
    xy = np.column_stack(((uv[:, 0] - W / 2) * metres_per_px, -(uv[:, 1] - H / 2) * metres_per_px))
    xy += np.asarray(world_offset)
    v_true = (shift_px[0] * metres_per_px, -shift_px[1] * metres_per_px, 0.0)
    s = velocity_sigma
    if kind == "cartesian":
        motion = dict(kind="cartesian", dem=0.0, dem_sigma=0.0, xy_sigma=(0.1, 0.1), vxyz=v_true,
                      vxyz_sigma=(s, s, 0.0), axyz=(0, 0, 0), axyz_sigma=(0.05, 0.05, 0.0))
    elif kind == "tangent_cartesian":
        motion = dict(kind=kind, dem=0.0, dem_sigma=0.3, xy_sigma=(0.1, 0.1), vxy=v_true[:2], vxy_sigma=(s, s),
                      axy=(0, 0), axy_sigma=(0.05, 0.05), slope_sigma=0.1)
    elif kind == "tangent_cylindrical":
        motion = dict(kind=kind, dem=0.0, dem_sigma=0.3, xy_sigma=(0.1, 0.1),
                      vrth=(float(np.hypot(v_true[0], v_true[1])), float(np.arctan2(v_true[1], v_true[0]))),
                      vrth_sigma=(s, 0.3), arth=(0, 0), arth_sigma=(0.05, 0.02), slope_sigma=0.1)
    else:
        speed = float(np.hypot(v_true[0], v_true[1]))
        theta = float(np.arctan2(v_true[1], v_true[0]))
        motion = dict(kind="cylindrical", dem=0.0, dem_sigma=1.0, xy_sigma=(0.1, 0.1), vrthz=(speed, theta, 0.0),
                      vrthz_sigma=(s, 0.3, 0.05), arthz=(0, 0, 0), arthz_sigma=(0.05, 0.02, 0.01))
    dts = [T0 + t * DAY for t in range(n_frames)]
    return Scene(
        observers=[ObserverScene(frames, cams, dts)], points=xy, motion=motion, n_particles=n_particles,
        tile_size=tile_size, truth_velocity=v_true,
        meta=dict(seed=seed, imgsz=imgsz, shift_px=shift_px, metres_per_px=metres_per_px, height=height),
    )


def as_raster_frames(scene: Scene, observer: int = 0) -> Scene:
    """Turn one observer of a distortion-free nadir scene into a sequence of orthoimages (``Raster`` frames with a
    datetime): pixel (u, v) of the nadir camera is the ground cell at x = ox + (u - W/2) m, y = oy - (v - H/2) m."""
    obs = scene.observers[observer]
    m = scene.meta["metres_per_px"]
    for vec in obs.cams:
        W, H = (int(v) for v in vec[6:8])
        ox, oy = vec[0], vec[1]
        grid = np.zeros(20)
        grid[0:6] = (ox - W / 2 * m, oy + H / 2 * m, np.nan, ox + W / 2 * m, oy - H / 2 * m, np.nan)
        grid[6:8] = (W, H)
        vec[:] = grid
    return scene


def build(scene: Scene, api, points: Optional[Sequence[int]] = None):
    """Instantiate ``api.Camera/Image/Observer/<Motion>`` objects (``api`` = this package or the
    reference package).  Returns ``(observers, motion_models)``."""
    observers = []
    for o, obs in enumerate(scene.observers):
        images = []
        for i, (frame, vec, dt) in enumerate(zip(obs.frames, obs.cams, obs.datetimes)):
            if np.isnan(vec[2]):  # a raster frame: [xlim[0], ylim[0], NaN, xlim[1], ylim[1], NaN, nx, ny, ...] (as_raster_frames)
                images.append(api.Raster(frame, x=(vec[0], vec[3]), y=(vec[1], vec[4]), datetime=dt))
                continue
            cam = api.Camera(imgsz=tuple(int(v) for v in vec[6:8]), f=tuple(vec[8:10]), c=tuple(vec[10:12]),
                             k=tuple(vec[12:18]), p=tuple(vec[18:20]), xyz=tuple(vec[0:3]), viewdir=tuple(vec[3:6]))
            img = api.Image(f"obs{o}_frame{i}", cam=cam, datetime=dt)
            img.array = frame
            images.append(img)
        observers.append(api.Observer(images, sigma=obs.sigma))
    params = dict(scene.motion)
    kind = params.pop("kind")
    cls = {"cartesian": "CartesianMotion", "cylindrical": "CylindricalMotion", "tangent_cartesian": "TangentCartesianMotion",
           "tangent_cylindrical": "TangentCylindricalMotion"}[kind]
    cls = getattr(api, cls)
    for key in ("dem", "dem_sigma"):  # gridded surfaces travel as dict(array=, x=, y=) and become the api's Raster
        if isinstance(params.get(key), dict):
            params[key] = api.Raster(params[key]["array"], x=params[key]["x"], y=params[key]["y"])
    sel = range(len(scene.points)) if points is None else points
    models = [cls(xy=scene.points[i], time_unit=scene.time_unit, n=scene.n_particles, **params) for i in sel]
    return observers, models
