"""CPU oracle for the ``glimpse.Tracker`` particle-filter hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module,
and only as the checker / the timed CPU baseline.  ``glimpse_b200`` never imports it.

It is a functional NumPy restatement of the reference algorithm (ezwelty/glimpse,
``/root/reference``; citations are relative to ``/root/reference/src/glimpse``).  The three
third-party kernels the reference calls on this path are called here as well
(``scipy.ndimage.median_filter``, ``cv2.matchTemplate``, ``scipy.interpolate.RectBivariateSpline``),
and each also has an independent closed-form restatement (``exact`` variants) used to bound
the library noise.

Parity pin: ``tests/golden/*.npz`` hold outputs of the *reference itself* (imported in the build
container by ``tests/golden/make_golden.py`` through ``oracle/ref_shim.py``); ``tests/test_oracle.py``
checks this oracle against them, together with the known-answer vectors of the reference's own
tests/doctests (``tests/test_camera.py:34-88``, ``camera.py:615-620``, ``helpers.py:335-342,451-456,
482-487,827-829``).

All state is plain ``numpy``: a camera is its 20-vector ``[xyz, viewdir, imgsz, f, c, k1..k6, p1, p2]``
(``camera.py:101,127-198``), a particle set is ``(n, 6)`` float64 ``[x, y, z, vx, vy, vz]``.
"""
from __future__ import annotations

import math
import warnings
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# Camera model (camera.py:239-280, 591-663, 1138-1264, 1305-1337, 1435-1519)
# --------------------------------------------------------------------------------------


def rotation_matrix(viewdir_deg: Sequence[float]) -> np.ndarray:
    """World->camera rotation from (yaw, pitch, roll) in degrees (camera.py:239-280)."""
    yaw, pitch, roll = np.deg2rad(np.asarray(viewdir_deg, dtype=float))
    c1, c2, c3 = math.cos(yaw), math.cos(pitch), math.cos(roll)
    s1, s2, s3 = math.sin(yaw), math.sin(pitch), math.sin(roll)
    return np.array(
        [
            [c1 * c3 + s1 * s2 * s3, c1 * s2 * s3 - c3 * s1, -c2 * s3],
            [c3 * s1 * s2 - c1 * s3, s1 * s3 + c1 * c3 * s2, -c2 * c3],
            [c2 * s1, c1 * c2, s2],
        ]
    )


def _radial(k: np.ndarray, r2: np.ndarray) -> np.ndarray:
    """Rational radial multiplier, zero coefficients skipped (camera.py:1138-1163)."""
    num = 1
    if k[0]:
        num = num + k[0] * r2
    if k[1]:
        num = num + k[1] * r2 * r2
    if k[2]:
        num = num + k[2] * r2 * r2 * r2
    if k[3] or k[4] or k[5]:
        den = 1
        if k[3]:
            den = den + k[3] * r2
        if k[4]:
            den = den + k[4] * r2 * r2
        if k[5]:
            den = den + k[5] * r2 * r2 * r2
        num = num / den
    return num


def _tangential(p: np.ndarray, xy: np.ndarray, r2: np.ndarray) -> np.ndarray:
    """Tangential additive term (camera.py:1165-1178)."""
    cross = xy[:, 0] * xy[:, 1]
    tx = 2 * cross * p[0] + p[1] * (r2 + 2 * xy[:, 0] ** 2)
    ty = p[0] * (r2 + 2 * xy[:, 1] ** 2) + 2 * cross * p[1]
    return np.column_stack((tx, ty))


def distort(cam: np.ndarray, xy: np.ndarray) -> np.ndarray:
    """Camera coordinates -> distorted camera coordinates (camera.py:1180-1196)."""
    k, p = cam[12:18], cam[18:20]
    if not k.any() and not p.any():
        return xy
    out = xy.copy()
    r2 = (xy ** 2).sum(axis=1)
    if k.any():
        out *= np.asarray(_radial(k, r2))[:, None]
    if p.any():
        out += _tangential(p, xy, r2)
    return out


def undistort(cam: np.ndarray, xy: np.ndarray, iterations: int = 20) -> np.ndarray:
    """Inverse of :func:`distort` (camera.py:1198-1264, 1305-1337).

    k1-only -> closed-form cubic; otherwise 20 fixed-point (Oulu) iterations with the
    combined update (the reference's separate radial-only branch is unreachable).
    """
    k, p = cam[12:18], cam[18:20]
    if not k.any() and not p.any():
        return xy
    if k[0] and not k[1:].any() and not p.any():
        phi = np.arctan2(xy[:, 1], xy[:, 0])
        q = -1 / (3 * k[0])
        rr = -xy[:, 0] / (2 * k[0] * np.cos(phi))
        three = rr ** 2 < q ** 3
        r = np.full(len(xy), np.nan)
        if three.any():
            th = np.arccos(rr[three] * q ** -1.5)
            r[three] = -2 * np.sqrt(q) * np.cos((th - 2 * np.pi) / 3)
        one = ~three
        if one.any():
            a = -np.sign(rr[one]) * (np.abs(rr[one]) + np.sqrt(rr[one] ** 2 - q ** 3)) ** (1.0 / 3)
            b = np.zeros(a.shape)
            nz = a != 0
            b[nz] = q / a[nz]
            r[one] = a + b
        return np.column_stack((np.cos(phi), np.sin(phi))) * r[:, None]
    est = xy
    for _ in range(iterations):
        r2 = (est ** 2).sum(axis=1)
        if p.any() and not k.any():
            est = xy - _tangential(p, est, r2)
        else:
            est = (xy - _tangential(p, est, r2)) * (1 / np.asarray(_radial(k, r2)))[:, None]
    return est


def grid_vector(xlim: Sequence[float], ylim: Sequence[float], size: Sequence[int]) -> np.ndarray:
    """A raster frame (an orthoimage: ``Raster`` with a datetime as an Observer's image) in the 20 slots of a camera vector:
    ``[xlim[0], ylim[0], NaN, xlim[1], ylim[1], NaN, nx, ny, 0...]`` — the NaN camera height marks it."""
    out = np.zeros(20)
    out[0:6] = (xlim[0], ylim[0], np.nan, xlim[1], ylim[1], np.nan)
    out[6:8] = size
    return out


def is_grid_vector(cam: np.ndarray) -> bool:
    return bool(np.isnan(cam[2]))


def grid_cell_size(cam: np.ndarray) -> np.ndarray:
    """``Grid.d`` (raster.py:119-122): signed cell size from the outer limits and the size."""
    return np.hstack((np.diff(cam[[0, 3]]), np.diff(cam[[1, 4]]))) / cam[6:8].astype(int)


def project(cam: np.ndarray, xyz: np.ndarray, correction: Optional[Tuple[float, float]] = None, directions: bool = False) -> np.ndarray:
    """World -> image coordinates (camera.py:591-628, 1435-1470, 1499-1508).

    ``correction`` = (radius, refraction) enables the curvature/refraction term
    (helpers.py:1771-1790).  Points behind the camera give NaN.  ``directions``: ``xyz`` are ray directions from the
    camera (no translation, no correction: camera.py:1448-1449).
    """
    cam = np.asarray(cam, dtype=float)
    if is_grid_vector(cam):  # a raster frame: (xy - (xlim[0], ylim[0])) / d, z unused (raster.py:423-445)
        return (np.asarray(xyz, dtype=float)[:, 0:2] - (cam[0], cam[1])) / grid_cell_size(cam)
    d = np.asarray(xyz, dtype=float) if directions else np.asarray(xyz, dtype=float) - cam[0:3]
    if correction is not None and not directions:
        radius, refraction = correction
        d[:, 2] += (refraction - 1) * (d[:, 0:2] ** 2).sum(axis=1) / (2 * radius)
    R = rotation_matrix(cam[3:6])
    c = np.dot(d, R.T)
    with np.errstate(invalid="ignore", divide="ignore"):
        xy = c[:, 0:2] / c[:, 2:3]
    xy[c[:, 2] <= 0] = np.nan
    xy = distort(cam, xy)
    imgsz = cam[6:8].astype(int)
    return xy * cam[8:10] + (imgsz / 2 + cam[10:12])


def unproject(cam: np.ndarray, uv: np.ndarray, directions: bool = True, depth=1) -> np.ndarray:
    """Image -> world ray directions / points at depth (camera.py:630-663, 1472-1497, 1510-1519)."""
    cam = np.asarray(cam, dtype=float)
    imgsz = cam[6:8].astype(int)
    xy = (np.asarray(uv, dtype=float) - (imgsz * 0.5 + cam[10:12])) * (1 / cam[8:10])
    xy = undistort(cam, xy)
    R = rotation_matrix(cam[3:6])
    xyz = np.dot(xy, R[0:2, :])
    xyz += R.T[:, 2]
    if not isinstance(depth, (int, float)) or depth != 1:
        xyz *= np.atleast_1d(depth).reshape(-1, 1)
    if not directions:
        xyz += cam[0:3]
    return xyz


def _grid_intervals(x: np.ndarray, n: int):
    """Cell index and normalised distance of ``x`` on the pixel-centre grid 0.5, 1.5, .., n - 0.5 the way
    scipy.interpolate.RegularGridInterpolator finds them (``find_indices``: grid[i] <= x < grid[i + 1], clipped to
    [0, n - 2]; NaN stays NaN in the distance)."""
    with np.errstate(invalid="ignore"):
        i = np.floor(x - 0.5)
    i = np.where(np.isnan(i), 0, i).astype(np.int64)
    i = np.clip(i, 0, n - 2)
    return i, (x - (i + 0.5)) / 1.0


def project_image(frame: np.ndarray, src_cam: np.ndarray, dst_cam: np.ndarray, method: str = "linear") -> np.ndarray:
    """``Image.project`` (image.py:301-361): the frame resampled into another camera at the same position.

    Every pixel centre of the target camera is cast out as a ray (``uv_to_xyz``), projected into the source camera
    (``xyz_to_uv(directions=True)``) and the source band sampled there with ``scipy.interpolate.RegularGridInterpolator``
    on the pixel-centre grid (``bounds_error=False``: NaN outside), band by band; the result takes the frame's dtype
    (NaN becomes 0 in an integer frame).  The interpolator's arithmetic, restated from scipy 1.18 (the container's; the
    reference pins 1.4.1, whose linear form is the generic one below for every dtype):
      integer and float64 bands (integers are converted to float64): ``v00 * (1 - y0) * (1 - y1) + v01 * (1 - y0) * y1
      + v10 * y0 * (1 - y1) + v11 * y0 * y1`` left to right (``_rgi_cython.evaluate_linear_2d``);
      other floating bands: ``sum(v * ((1 * wy) * wx))`` over the same four corners in the same order (``_evaluate_linear``);
      nearest: the lower index where the normalised distance is <= 0.5, else the upper (``_evaluate_nearest``)."""
    src_cam, dst_cam = np.asarray(src_cam, dtype=float), np.asarray(dst_cam, dtype=float)
    if not all(src_cam[0:3] == dst_cam[0:3]):
        raise ValueError("Source and target cameras have different positions ('xyz')")
    W, H = (int(v) for v in dst_cam[6:8])
    sw, sh = (int(v) for v in src_cam[6:8])
    u = np.linspace(0.5, W - 0.5, W)
    v = np.linspace(0.5, H - 0.5, H)
    U, V = np.meshgrid(u, v)
    uv = np.column_stack((U.flatten(), V.flatten()))
    pvu = np.fliplr(project(src_cam, unproject(dst_cam, uv), directions=True))
    array = frame if frame.ndim == 3 else frame[:, :, None]
    out = np.empty((H, W, array.shape[2]), dtype=array.dtype)
    pv, pu = pvu[:, 0], pvu[:, 1]
    with np.errstate(invalid="ignore"):
        outside = (pv < 0.5) | (pv > sh - 0.5) | (pu < 0.5) | (pu > sw - 0.5)
    nans = np.isnan(pv) | np.isnan(pu)
    i0, y0 = _grid_intervals(pv, sh)
    i1, y1 = _grid_intervals(pu, sw)
    for b in range(array.shape[2]):
        band = array[:, :, b]
        if not np.issubdtype(band.dtype, np.inexact):
            band = band.astype(float)
        if method == "nearest":
            with np.errstate(invalid="ignore"):
                r = np.where(y0 <= 0.5, i0, i0 + 1)
                c = np.where(y1 <= 0.5, i1, i1 + 1)
            val = band[r, c].astype(float)
        elif method == "linear":
            if band.dtype == np.float64:
                val = (band[i0, i1] * (1 - y0) * (1 - y1) + band[i0, i1 + 1] * (1 - y0) * y1
                       + band[i0 + 1, i1] * y0 * (1 - y1) + band[i0 + 1, i1 + 1] * y0 * y1)
            else:
                val = np.zeros(len(pv))
                for (r, wy), (c, wx) in (((i0, 1 - y0), (i1, 1 - y1)), ((i0, 1 - y0), (i1 + 1, y1)),
                                         ((i0 + 1, y0), (i1, 1 - y1)), ((i0 + 1, y0), (i1 + 1, y1))):
                    val = val + band[r, c] * ((1.0 * wy) * wx)
        else:
            raise ValueError(f"Method '{method}' is not defined")
        val = np.where(outside | nans, np.nan, val)
        with np.errstate(invalid="ignore"):
            out[:, :, b] = val.reshape(H, W)
    return out


def grid_centres(lim: Sequence[float], n: int) -> np.ndarray:
    """``Grid.x`` / ``Grid.y`` (raster.py:139-174): cell centres in array order from the outer limits (first, last)."""
    lo, hi = min(lim), max(lim)
    half = abs((lim[1] - lim[0]) / n) / 2
    centres = np.linspace(start=lo + half, stop=hi - half, num=n)
    return centres[::-1] if lim[1] < lim[0] else centres


def viewshed(z: np.ndarray, xlim: Sequence[float], ylim: Sequence[float], origin: Sequence[float],
             correction: Optional[Tuple[float, float]] = None) -> np.ndarray:
    """``Raster.viewshed`` (raster.py:1293-1389): cells of the (ny, nx) surface ``z`` visible from ``origin`` (x, y, z).

    Cells are binned into rings by their rounded distance in cells and swept outwards: within a ring every cell's
    elevation ratio dz / distance is compared with the highest ratio seen so far along its heading, which is the
    previous ring's running maximum linearly interpolated (periodically) at the cell's heading.  Cells closer than half a
    cell to the origin (ring 0) are skipped when other rings exist — the reference starts at the first ring boundary —
    and a raster that is one ring only is all visible if that ring is ring 0.  ``correction`` = (radius, refraction)."""
    z = np.asarray(z, dtype=float)
    ny, nx = z.shape
    x, y = grid_centres(xlim, nx), grid_centres(ylim, ny)
    dx = np.tile(x - origin[0], ny)
    dy = np.repeat(y - origin[1], nx)
    dz = z.ravel() - origin[2]
    d2 = dx ** 2 + dy ** 2
    if correction is not None:
        radius, refraction = correction
        dz = dz + (refraction - 1) * d2 / (2 * radius)
    dist = np.sqrt(d2)
    cell = abs((xlim[1] - xlim[0]) / nx)
    ring = (dist * (1 / cell) + 0.5).astype(int)
    heading = np.arctan2(dy, dx)
    order = np.lexsort((heading, ring))
    ring_sorted = ring[order]
    starts = np.flatnonzero(ring_sorted[1:] != ring_sorted[:-1]) + 1  # first sorted position of every ring but the first
    if ring_sorted[0] != 0:
        starts = np.concatenate(([0], starts))
    elif len(starts) == 0:
        return np.ones(z.shape, dtype=bool)
    bounds = np.concatenate((starts, [len(order)]))
    first = order[bounds[0]:bounds[1]]
    dist[first[dist[first] == 0]] = np.nan
    with np.errstate(invalid="ignore", divide="ignore"):
        ratio = dz / dist
    visible = np.zeros(z.size, dtype=bool)
    horizon_heading = horizon = None
    horizon_has_nan = False
    for k in range(len(bounds) - 1):
        cells = order[bounds[k]:bounds[k + 1]]
        h, r = heading[cells], ratio[cells]
        if k == 0:
            seen = ~np.isnan(r)
            horizon = r
            horizon_has_nan = bool(np.isnan(r).any())
        else:
            horizon = np.interp(h, horizon_heading, horizon, period=2 * np.pi)
            with np.errstate(invalid="ignore"):
                seen = r > horizon
            if horizon_has_nan:
                unknown = np.isnan(horizon)
                opened = unknown & ~np.isnan(r)
                seen |= opened
                if np.count_nonzero(unknown) == np.count_nonzero(opened):
                    horizon_has_nan = False
            horizon[seen] = r[seen]
        visible[cells] = seen
        horizon_heading = h
    return visible.reshape(z.shape)


def inframe(cam: np.ndarray, uv: np.ndarray) -> np.ndarray:
    """(camera.py:700-718)."""
    imgsz = np.asarray(cam)[6:8].astype(int)
    with np.errstate(invalid="ignore"):
        return np.all((uv >= 0) & (uv <= imgsz), axis=1)


# --------------------------------------------------------------------------------------
# Surfaces: DEM / DEM sigma / viewshed (raster.py:313-341, 891-1027)
# --------------------------------------------------------------------------------------


@dataclass
class Surface:
    """Constant (0-D) or gridded (2-D) raster sampled in point mode.

    ``array`` is (ny, nx); ``xlim``/``ylim`` are the outer limits in array order (left,right)/(top,bottom)
    exactly as the reference's ``Raster(array, x=xlim, y=ylim)``.  A scalar ``array`` is the
    reference's 0-D raster with infinite limits (motion.py:136-141).
    """

    array: np.ndarray
    xlim: Tuple[float, float] = (-np.inf, np.inf)
    ylim: Tuple[float, float] = (-np.inf, np.inf)

    def __post_init__(self):
        self.array = np.asarray(self.array, dtype=float)
        self.constant = self.array.ndim == 0 or self.array.size == 1

    def _centers(self):
        ny, nx = self.array.shape
        dx = (self.xlim[1] - self.xlim[0]) / nx
        dy = (self.ylim[1] - self.ylim[0]) / ny
        xs = np.linspace(min(self.xlim) + abs(dx) / 2, max(self.xlim) - abs(dx) / 2, nx)
        ys = np.linspace(min(self.ylim) + abs(dy) / 2, max(self.ylim) - abs(dy) / 2, ny)
        z = self.array.T  # (nx, ny), indexed by increasing array order
        if dx < 0:
            z = z[::-1, :]
        if dy < 0:
            z = z[:, ::-1]
        return xs, ys, z

    def sample(self, xy: np.ndarray, order: int = 1) -> np.ndarray:
        """Point-mode ``Raster.sample`` with ``bounds_error=True`` (raster.py:913-1027)."""
        xy = np.asarray(xy, dtype=float)
        lo = np.array((min(self.xlim), min(self.ylim)))
        hi = np.array((max(self.xlim), max(self.ylim)))
        if not np.all((xy >= lo) & (xy <= hi)):
            raise ValueError("Some of the sampling coordinates are out of bounds")
        if self.constant:
            return np.full(len(xy), self.array.flat[0])
        xs, ys, z = self._centers()
        out = np.empty(len(xy))

        def locate(grid, v):
            i = np.clip(np.searchsorted(grid, v) - 1, 0, len(grid) - 2)
            t = (v - grid[i]) / (grid[i + 1] - grid[i])
            return i, t

        ix, tx = locate(xs, xy[:, 0])
        iy, ty = locate(ys, xy[:, 1])
        if order == 0:
            jx = np.where(tx <= 0.5, ix, ix + 1)
            jy = np.where(ty <= 0.5, iy, iy + 1)
            return z[jx, jy]
        out = (
            z[ix, iy] * (1 - tx) * (1 - ty)
            + z[ix, iy + 1] * (1 - tx) * ty
            + z[ix + 1, iy] * tx * (1 - ty)
            + z[ix + 1, iy + 1] * tx * ty
        )
        return out


# --------------------------------------------------------------------------------------
# Motion models (track/motion.py:92-311)
# --------------------------------------------------------------------------------------


@dataclass
class MotionSpec:
    """Parameters of CartesianMotion (kind='cartesian', motion.py:92-204), CylindricalMotion
    (kind='cylindrical', motion.py:207-311), TangentCartesianMotion (kind='tangent_cartesian',
    motion.py:314-420) or TangentCylindricalMotion (kind='tangent_cylindrical', motion.py:423-522).
    For the cylindrical models ``v``/``a`` hold (d radius/dt, theta[, dz/dt]) and
    (d2 radius/dt2, d theta/dt[, d2z/dt2]); the tangent models use the first two components only."""

    xy: Sequence[float]
    n: int = 1000
    kind: str = "cartesian"
    dem: Surface = field(default_factory=lambda: Surface(0.0))
    dem_sigma: Surface = field(default_factory=lambda: Surface(0.0))
    xy_sigma: Sequence[float] = (0, 0)
    v: Sequence[float] = (0, 0, 0)
    v_sigma: Sequence[float] = (0, 0, 0)
    a: Sequence[float] = (0, 0, 0)
    a_sigma: Sequence[float] = (0, 0, 0)
    slope_sigma: float = 0.0

    @property
    def tangent(self) -> bool:
        return self.kind.startswith("tangent")

    @property
    def cylindrical(self) -> bool:
        return self.kind.endswith("cylindrical")


def init_particles(m: MotionSpec, randn: Callable = np.random.randn) -> np.ndarray:
    """Draw order: randn(n,2), randn(n), then randn(n,3) (motion.py:149-163, 260-283) or, for the tangent
    models, randn(n,2) — their vz stays 0 (motion.py:378-390, 485-505)."""
    ps = np.zeros((m.n, 6))
    ps[:, 0:2] = np.asarray(m.xy, float) + np.asarray(m.xy_sigma, float) * randn(m.n, 2)
    if m.tangent:
        z_offsets = m.dem_sigma.sample(ps[:, 0:2]) * randn(m.n)
        ps[:, 2] = m.dem.sample(ps[:, 0:2]) + z_offsets
        vel = np.asarray(m.v, float)[:2] + np.asarray(m.v_sigma, float)[:2] * randn(m.n, 2)
        if m.cylindrical:
            vel = np.column_stack((vel[:, 0] * np.cos(vel[:, 1]), vel[:, 0] * np.sin(vel[:, 1])))
        ps[:, 3:5] = vel
        return ps
    ps[:, 2] = m.dem.sample(ps[:, 0:2])
    ps[:, 2] += m.dem_sigma.sample(ps[:, 0:2]) * randn(m.n)
    vel = np.asarray(m.v, float) + np.asarray(m.v_sigma, float) * randn(m.n, 3)
    if m.cylindrical:
        vel = np.column_stack((vel[:, 0] * np.cos(vel[:, 1]), vel[:, 0] * np.sin(vel[:, 1]), vel[:, 2]))
    ps[:, 3:6] = vel
    return ps


def evolve_particles(m: MotionSpec, ps: np.ndarray, tau: float, randn: Callable = np.random.randn) -> None:
    """In-place random-acceleration step over ``tau`` time units (motion.py:165-179, 285-311); the tangent
    models move in x, y only and carry each particle's offset above the DEM along, widened by a random walk
    proportional to the distance travelled (motion.py:392-420, 507-522; draws: randn(n,2) then randn(n))."""
    n = len(ps)
    if m.tangent:
        acc = np.asarray(m.a, float)[:2] + np.asarray(m.a_sigma, float)[:2] * randn(n, 2)
        if m.cylindrical:
            vx, vy = ps[:, 3], ps[:, 4]
            with np.errstate(invalid="ignore", divide="ignore"):
                speed = np.sqrt(vx ** 2 + vy ** 2)
                acc = np.column_stack((acc[:, 0] * (vx / speed) - vy * acc[:, 1], acc[:, 0] * (vy / speed) + vx * acc[:, 1]))
        dxy = tau * ps[:, 3:5] + 0.5 * acc * tau ** 2
        z_offsets = ps[:, 2] - m.dem.sample(ps[:, 0:2])
        z_offsets += m.slope_sigma * randn(n) * (dxy ** 2).sum(axis=1) ** 0.5
        ps[:, 0:2] += dxy
        ps[:, 2] = m.dem.sample(ps[:, 0:2]) + z_offsets
        ps[:, 3:5] += tau * acc
        return
    acc = np.asarray(m.a, float) + np.asarray(m.a_sigma, float) * randn(n, 3)
    if m.cylindrical:
        vx, vy = ps[:, 3], ps[:, 4]
        with np.errstate(invalid="ignore", divide="ignore"):
            speed = np.sqrt(vx ** 2 + vy ** 2)
            acc = np.column_stack(
                (acc[:, 0] * (vx / speed) - vy * acc[:, 1], acc[:, 0] * (vy / speed) + vx * acc[:, 1], acc[:, 2])
            )
    ps[:, 0:3] += tau * ps[:, 3:6] + 0.5 * acc * tau ** 2
    ps[:, 3:6] += tau * acc


def surface_log_likelihood(m: MotionSpec, ps: np.ndarray) -> Optional[np.ndarray]:
    """(dem(xy) - z)^2 / (2 sigma^2) where sigma != 0 (motion.py:181-204); the tangent models inherit the
    base class's ``None`` (motion.py:77-89)."""
    if m.tangent:
        return None
    z = m.dem.sample(ps[:, 0:2])
    zs = m.dem_sigma.sample(ps[:, 0:2])
    nz = np.nonzero(zs)[0]
    ll = np.zeros(len(ps))
    ll[nz] = 1 / (2 * zs[nz] ** 2) * (z[nz] - ps[nz, 2]) ** 2
    return ll


# --------------------------------------------------------------------------------------
# Tiles (tracker.py:494-561; observer.py:115-144; raster.py:343-421; helpers.py:324-344,433-493)
# --------------------------------------------------------------------------------------


def snap_tile_box(uv: np.ndarray, size: Sequence[int], imgsz: Sequence[int]) -> np.ndarray:
    """Integer (left, top, right, bottom) template box around ``uv`` (observer.py:115-130,
    raster.py:343-421): corners rounded half-up; IndexError if the unsnapped box leaves the image."""
    half = np.multiply(size, 0.5)
    corners = np.vstack((uv - half, uv + half))
    if not np.all((corners >= 0) & (corners <= np.asarray(imgsz))):
        raise IndexError("Box extends beyond grid bounds")
    return np.floor(corners + 0.5).flatten().astype(int)


def to_gray(tile: np.ndarray) -> np.ndarray:
    """(tracker.py:523-524)."""
    return tile.mean(axis=2) if tile.ndim > 2 else tile


def normalize(a: np.ndarray) -> np.ndarray:
    """(helpers.py:324-344)."""
    return (a - a.mean()) * (1 / a.std())


def value_cdf(a: np.ndarray):
    """Sorted unique values and P(value <= v) (helpers.py:433-464)."""
    values, counts = np.unique(a, return_counts=True)
    return values, np.cumsum(counts) / a.size


def cdf_match(a: np.ndarray, cdf) -> np.ndarray:
    """(helpers.py:467-493)."""
    _, inverse, counts = np.unique(a, return_inverse=True, return_counts=True)
    q = np.cumsum(counts) / a.size
    mapped = np.interp(q, cdf[1], cdf[0])
    return mapped[inverse.ravel()].reshape(a.shape)


def median_window(tile: np.ndarray, size=(5, 5), exact: bool = False, **filter_kwargs) -> np.ndarray:
    """``scipy.ndimage.median_filter(tile, size, mode='reflect')`` or, with ``exact``, a direct
    restatement (edge pixel duplicated; rank ``n // 2`` of the window).  ``filter_kwargs`` = the other entries of
    ``Tracker.highpass`` (``mode``, ``cval``, ``origin``): always scipy's own filter."""
    if not exact or filter_kwargs:
        import scipy.ndimage

        return scipy.ndimage.median_filter(tile, size=size, **filter_kwargs)
    hy, hx = size[0] // 2, size[1] // 2
    padded = np.pad(tile, ((hy, size[0] - 1 - hy), (hx, size[1] - 1 - hx)), mode="symmetric")
    windows = np.lib.stride_tricks.sliding_window_view(padded, size).reshape(tile.shape + (-1,))
    k = (size[0] * size[1]) // 2
    return np.partition(windows, k, axis=2)[:, :, k]


def prepare_tile(pixels: np.ndarray, histogram=None, size=(5, 5), exact_median: bool = False, **filter_kwargs):
    """gray -> z-score -> (search: CDF match) -> (template: CDF) -> minus 5x5 median (tracker.py:522-534).
    Returns (tile, cdf of the pre-filter tile)."""
    tile = normalize(to_gray(pixels))
    if histogram is not None:
        tile = cdf_match(tile, histogram)
    cdf = value_cdf(tile)
    tile = tile - median_window(tile, size=size, exact=exact_median, **filter_kwargs)
    return tile, cdf


def search_window(uv: np.ndarray, size: Sequence[int], imgsz: Sequence[int], kx: int = 3, ky: int = 3):
    """Integer box around all projected particles grown by half a template, widened so the SSE
    surface has > k cells; None if it leaves the image (tracker.py:576-601)."""
    size = np.asarray(size)
    half = size * 0.5
    box = np.vstack((uv.min(axis=0) - half, uv.max(axis=0) + half))
    ncols = ky - ((box[1, 0] - box[0, 0]) - size[0])
    if ncols > 0:
        box[:, 0] += np.array((-ncols, ncols)) * 0.5
    nrows = kx - ((box[1, 1] - box[0, 1]) - size[1])
    if nrows > 0:
        box[:, 1] += np.array((-nrows, nrows)) * 0.5
    with np.errstate(invalid="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ibox = np.vstack((np.floor(box[0]), np.ceil(box[1]))).astype(int)
    if not np.isfinite(box).all():
        return None
    if not np.all((ibox >= 0) & (ibox <= np.asarray(imgsz))):
        return None
    return ibox.ravel()


def ssd_surface(search: np.ndarray, template: np.ndarray, exact: bool = False) -> np.ndarray:
    """Area-normalised sum of squared differences, float32 (tracker.py:609-614).

    ``exact=False``: ``cv2.matchTemplate(TM_SQDIFF)`` as the reference calls it.
    ``exact=True``: direct float64 sum over the float32-rounded tiles, rounded once to float32."""
    s32 = search.astype(np.float32)
    t32 = template.astype(np.float32)
    h, w = t32.shape
    if exact:
        win = np.lib.stride_tricks.sliding_window_view(s32.astype(np.float64), (h, w))
        sse = ((win - t32.astype(np.float64)) ** 2).sum(axis=(2, 3)).astype(np.float32)
    else:
        import cv2

        sse = cv2.matchTemplate(s32, templ=t32, method=cv2.TM_SQDIFF)
    size = np.array((w, h))
    sse *= 1 / (size[0] * size[1])
    return sse


def surface_box(box: np.ndarray, size: Sequence[int], duv: np.ndarray) -> np.ndarray:
    """Geo-reference of the SSE surface (tracker.py:615-620)."""
    edge = np.asarray(size) * 0.5 - 0.5
    out = box + np.concatenate((edge, -edge))
    out += np.tile(duv, 2)
    return out


def notaknot_slopes(y: np.ndarray) -> np.ndarray:
    """Node derivatives of the not-a-knot cubic spline through unit-spaced samples along axis 0
    (m >= 4).  Tridiagonal system: s[i-1] + 4 s[i] + s[i+1] = 3 (y[i+1] - y[i-1]) inside,
    s[0] + 2 s[1] = (5 d0 + d1) / 2 and its mirror at the ends."""
    m = y.shape[0]
    rhs = np.empty_like(y, dtype=float)
    rhs[1:-1] = 3 * (y[2:] - y[:-2])
    rhs[0] = (5 * (y[1] - y[0]) + (y[2] - y[1])) / 2
    rhs[-1] = (5 * (y[-1] - y[-2]) + (y[-2] - y[-3])) / 2
    A = np.zeros((m, m))
    for i in range(1, m - 1):
        A[i, i - 1], A[i, i], A[i, i + 1] = 1, 4, 1
    A[0, 0], A[0, 1] = 1, 2
    A[-1, -2], A[-1, -1] = 2, 1
    return np.linalg.solve(A, rhs.reshape(m, -1)).reshape(y.shape)


def fitpack_interpolation_knots(x: np.ndarray, k: int) -> np.ndarray:
    """Knots of FITPACK's interpolating spline (s = 0) of degree k through the sites x (fpregr.f / fpcurf.f): k + 1 copies
    of the end sites and m - k - 1 interior knots — the sites x[(k+1)/2 .. m-1-(k+1)/2] for an odd degree, the midpoints
    (x[j-1] + x[j]) / 2, j = k/2+1 .. m-1-k/2, for an even one."""
    m = len(x)
    if k % 2:
        interior = x[(k + 1) // 2: m - (k + 1) // 2]
    else:
        j = np.arange(k // 2 + 1, m - k // 2)
        interior = (x[j - 1] + x[j]) * 0.5
    return np.concatenate((np.repeat(x[0], k + 1), interior, np.repeat(x[-1], k + 1)))


def _bspline_sample(uv, F, cu, cv, kx, ky):
    """Tensor-product interpolating spline of degrees (kx along rows, ky along columns) restated with SciPy's B-spline
    collocation solver on FITPACK's knots (independent of FITPACK's own regrid / bispev), evaluated at clamped arguments."""
    import scipy.interpolate

    tv, tu = fitpack_interpolation_knots(cv, kx), fitpack_interpolation_knots(cu, ky)
    rows = scipy.interpolate.make_interp_spline(cu, F.T, k=ky, t=tu)      # along the columns, for every row: c[(nu), (nv)]
    coef = scipy.interpolate.make_interp_spline(cv, rows.c.T, k=kx, t=tv)  # then along the rows: c[(nv), (nu)]
    u = np.clip(uv[:, 0], cu[0], cu[-1])
    v = np.clip(uv[:, 1], cv[0], cv[-1])
    bu = scipy.interpolate.BSpline.design_matrix(u, tu, ky).toarray()  # (n, nu)
    bv = scipy.interpolate.BSpline.design_matrix(v, tv, kx).toarray()  # (n, nv)
    return np.einsum("nv,vu,nu->n", bv, coef.c, bu)


def spline_sample(uv: np.ndarray, surface: np.ndarray, box: np.ndarray, exact: bool = False, kx: int = 3, ky: int = 3) -> np.ndarray:
    """Sample ``surface`` (cell centres inside ``box``) at ``uv`` with the interpolating spline of degree ``kx``
    along the rows (v) and ``ky`` along the columns (u) (observer.py:178-214; ``Tracker.interpolation``).
    ``exact=False``: FITPACK via RectBivariateSpline(s=0).
    ``exact=True`` (degrees 1 and 3 only): tensor product of not-a-knot cubics (Hermite form) and / or piecewise-linear
    interpolants — FITPACK's degree-1 interpolating spline has a knot at every data site — with FITPACK's clamped evaluation."""
    lo, hi = box[0:2], box[2:4]
    if not np.all((uv >= lo) & (uv <= hi)):
        raise ValueError("Some sampling points are outside box")
    du = (box[2] - box[0]) / surface.shape[1]
    dv = (box[3] - box[1]) / surface.shape[0]
    cu = np.arange(box[0] + du * 0.5, box[2])
    cv = np.arange(box[1] + dv * 0.5, box[3])
    if not exact:
        import scipy.interpolate

        f = scipy.interpolate.RectBivariateSpline(cv, cu, surface, kx=kx, ky=ky)
        return f(uv[:, 1], uv[:, 0], grid=False)
    if kx not in (1, 3) or ky not in (1, 3):
        return _bspline_sample(uv, surface.astype(float), cu, cv, kx, ky)
    F = surface.astype(float)
    zero = np.zeros_like(F)
    Fv = notaknot_slopes(F) if kx == 3 else zero  # d/dv along rows axis
    Fu = notaknot_slopes(F.T).T if ky == 3 else zero
    Fuv = notaknot_slopes(Fu) if kx == 3 and ky == 3 else zero
    mv, mu = F.shape
    x = np.clip(uv[:, 0], cu[0], cu[-1]) - cu[0]
    y = np.clip(uv[:, 1], cv[0], cv[-1]) - cv[0]
    j = np.minimum(np.floor(x).astype(int), mu - 2)
    i = np.minimum(np.floor(y).astype(int), mv - 2)
    tx, ty = x - j, y - i

    def basis(t, k):
        if k == 1:
            return 1 - t, 0 * t, t, 0 * t
        t2, t3 = t * t, t * t * t
        return 2 * t3 - 3 * t2 + 1, t3 - 2 * t2 + t, -2 * t3 + 3 * t2, t3 - t2

    a0, a1, a2, a3 = basis(tx, ky)  # value0, slope0, value1, slope1 along u
    b0, b1, b2, b3 = basis(ty, kx)
    out = np.zeros(len(uv))
    for bi, di in ((b0, 0), (b2, 1)):
        for aj, dj in ((a0, 0), (a2, 1)):
            out += bi * aj * F[i + di, j + dj]
    for bi, di in ((b0, 0), (b2, 1)):
        for aj, dj in ((a1, 0), (a3, 1)):
            out += bi * aj * Fu[i + di, j + dj]
    for bi, di in ((b1, 0), (b3, 1)):
        for aj, dj in ((a0, 0), (a2, 1)):
            out += bi * aj * Fv[i + di, j + dj]
    for bi, di in ((b1, 0), (b3, 1)):
        for aj, dj in ((a1, 0), (a3, 1)):
            out += bi * aj * Fuv[i + di, j + dj]
    return out


# --------------------------------------------------------------------------------------
# Weights, resampling, moments (tracker.py:72-104, 126-223)
# --------------------------------------------------------------------------------------


def weights_from_log_likelihoods(terms: List[np.ndarray]) -> np.ndarray:
    """(tracker.py:145-149)."""
    return np.exp(-sum(terms)) + 1e-300


def systematic_indices(weights: np.ndarray, u: float) -> np.ndarray:
    """(tracker.py:168-176)."""
    n = len(weights)
    wn = weights / weights.sum()
    positions = (np.arange(n) + u) * (1 / n)
    return np.searchsorted(np.cumsum(wn), positions)


def stratified_indices(weights: np.ndarray, u: np.ndarray) -> np.ndarray:
    """(tracker.py:178-186): one uniform per stratum instead of one for all."""
    n = len(weights)
    wn = weights / weights.sum()
    positions = (np.arange(n) + u) * (1 / n)
    return np.searchsorted(np.cumsum(wn), positions)


def residual_indices(weights: np.ndarray, random: Callable = np.random.random):
    """tracker.py:188-203, statement by statement (the residuals are those of the normalised weights, and np.searchsorted runs
    on their — not monotone — cumulative sum).  Returns (indices, the uniforms drawn)."""
    n = len(weights)
    weights = weights / weights.sum()
    repetitions = (n * weights).astype(int)
    initial_indexes = np.repeat(np.arange(n), repetitions)
    residuals = weights - repetitions
    residuals *= 1 / residuals.sum()
    cumulative_sum = np.cumsum(residuals)
    cumulative_sum[-1] = 1.0
    u = random(n - len(initial_indexes))
    additional_indexes = np.searchsorted(cumulative_sum, u)
    return np.hstack((initial_indexes, additional_indexes)), u


def choice_indices(weights: np.ndarray, u: np.ndarray) -> np.ndarray:
    """(tracker.py:205-209): ``np.random.choice(arange(n), n, replace=True, p=w / w.sum())`` of the legacy generator =
    inverse-CDF sampling with ``random_sample(n)`` (numpy/random/mtrand.pyx, ``RandomState.choice``)."""
    cdf = (weights / weights.sum()).cumsum()
    cdf /= cdf[-1]
    return cdf.searchsorted(u, side="right")


def weighted_mean(ps: np.ndarray, w: np.ndarray) -> np.ndarray:
    return np.average(ps, weights=w, axis=0)


def weighted_sigma(ps: np.ndarray, w: np.ndarray, mean: np.ndarray) -> np.ndarray:
    return np.sqrt(np.average((ps - mean) ** 2, weights=w, axis=0))


def weighted_covariance(ps: np.ndarray, w: np.ndarray) -> np.ndarray:
    return np.cov(ps.T, aweights=w, ddof=0)


# --------------------------------------------------------------------------------------
# The filter loop (tracker.py:225-417)
# --------------------------------------------------------------------------------------


@dataclass
class ObserverSpec:
    """One camera station: frames[i] is (H, W) or (H, W, C); cams[i] its 20-vector."""

    frames: List[np.ndarray]
    cams: np.ndarray
    sigma: float = 0.3
    corrections: Optional[List[Optional[Tuple[float, float]]]] = None

    def correction(self, i):
        return None if self.corrections is None else self.corrections[i]


@dataclass
class TrackResult:
    means: np.ndarray
    sigmas: np.ndarray
    errors: List[Optional[BaseException]]
    skipped: np.ndarray
    particles: Optional[np.ndarray] = None
    weights: Optional[np.ndarray] = None
    trace: Optional[List[Dict]] = None
    templates: Optional[List[Dict]] = None  # with trace=True: every template in creation order (point-major)


def track(
    observers: List[ObserverSpec],
    models: List[MotionSpec],
    taus: np.ndarray,
    image_index: np.ndarray,
    tile_size=(15, 15),
    observer_mask: Optional[np.ndarray] = None,
    viewshed: Optional[Surface] = None,
    return_covariances: bool = False,
    return_particles: bool = False,
    exact: bool = False,
    randn: Callable = None,
    random: Callable = None,
    resample_method: str = "systematic",
    trace: bool = False,
    raise_errors: bool = False,
    highpass_size=(5, 5),
    highpass_kwargs: Optional[Dict] = None,
    kx: int = 3,
    ky: int = 3,
) -> TrackResult:
    """Run the filter for every model (tracker.py:225-417, inner ``process`` 305-374).

    ``taus[i]`` = (datetimes[i+1] - datetimes[i]) / time_unit; ``image_index[t, o]`` = image of
    observer ``o`` matched to time ``t`` or -1 (tracker.py:466-492).  Draws come from the legacy
    global NumPy generator in the reference's order unless ``randn`` / ``random`` are supplied.
    ``exact`` switches the three library kernels to their closed-form restatements.
    ``highpass_size`` = ``Tracker.highpass["size"]`` as (rows, columns) or one integer (tracker.py:59, 530);
    ``highpass_kwargs`` = its other entries (``mode``, ``cval``, ``origin`` of ``scipy.ndimage.median_filter``).
    ``kx``, ``ky`` = ``Tracker.interpolation`` (tracker.py:60): spline degree along the rows / columns of the SSE surface.
    """
    if highpass_size is not None:  # (None: the window is given by highpass_kwargs['footprint'])
        if np.ndim(highpass_size) == 0:
            highpass_size = (int(highpass_size),) * 2
        highpass_size = tuple(int(v) for v in highpass_size)
    hp_kw = dict(highpass_kwargs or {})
    randn = randn or np.random.randn
    random = random or np.random.random
    T, O = image_index.shape
    P = len(models)
    if observer_mask is None:
        observer_mask = np.ones((P, O), dtype=bool)
    template_frame = (image_index >= 0).argmax(axis=0)
    means = np.full((P, T, 6), np.nan)
    sigmas = np.full((P, T, 6, 6) if return_covariances else (P, T, 6), np.nan)
    skipped = np.zeros((P, T, O), dtype=np.uint8)
    all_ps = np.full((P, T, models[0].n, 6), np.nan) if return_particles else None
    all_w = np.full((P, T, models[0].n), np.nan) if return_particles else None
    errors: List[Optional[BaseException]] = [None] * P
    log: List[Dict] = []
    made: List[Dict] = []
    for p, (model, mask) in enumerate(zip(models, observer_mask)):
        try:
            observed = (image_index[:, mask] >= 0).any(axis=1)
            first = int(np.argmax(observed))
            last = len(observed) - 1 - int(np.argmax(observed[::-1]))
            templates: List[Optional[Dict]] = [None] * O
            ps = w = None
            for t in range(first, last + 1):
                if t == first:
                    ps = init_particles(model, randn)
                else:
                    evolve_particles(model, ps, taus[t - 1], randn)
                if viewshed is not None and not all(viewshed.sample(ps[:, 0:2], order=0)):
                    raise ValueError("Some particles are on non-visible viewshed cells")
                if np.isnan(ps).any():
                    raise ValueError("Some particles have missing (NaN) values")
                if t == first:
                    w = np.ones(len(ps))
                for o in np.nonzero(mask & (template_frame == t))[0]:
                    obs, img = observers[o], int(image_index[t, o])
                    cam = obs.cams[img]
                    uv0 = project(cam, weighted_mean(ps, w)[None, 0:3], obs.correction(img)).ravel()
                    box = snap_tile_box(uv0, tile_size, cam[6:8].astype(int))
                    pixels = obs.frames[img][box[1]:box[3], box[0]:box[2]]
                    tile, cdf = prepare_tile(pixels, size=highpass_size, exact_median=exact, **hp_kw)
                    templates[o] = {"tile": tile, "cdf": cdf, "box": box, "duv": uv0 - box.reshape(2, -1).mean(axis=0)}
                    if trace:
                        made.append({"p": p, "obs": int(o), "img": img, "box": box, "tile": tile, "values": cdf[0], "quantiles": cdf[1]})
                step = {"p": p, "t": t} if trace else None
                if t > first:
                    terms = []
                    for o in range(O):
                        img = int(image_index[t, o]) if mask[o] else -1
                        if img < 0:
                            skipped[p, t, o] = 1
                            continue
                        obs, tpl = observers[o], templates[o]
                        cam = obs.cams[img]
                        size = tpl["tile"].shape[::-1]
                        uv = project(cam, ps[:, 0:3], obs.correction(img))
                        box = search_window(uv, size, cam[6:8].astype(int), kx=kx, ky=ky)
                        if box is None:
                            skipped[p, t, o] = 2
                            continue
                        pixels = obs.frames[img][box[1]:box[3], box[0]:box[2]]
                        search, _ = prepare_tile(pixels, histogram=tpl["cdf"], size=highpass_size, exact_median=exact, **hp_kw)
                        sse = ssd_surface(search, tpl["tile"], exact=exact)
                        sbox = surface_box(box, size, tpl["duv"])
                        sampled = spline_sample(uv, sse, sbox, exact=exact, kx=kx, ky=ky)
                        terms.append(sampled * (1 / (2 * obs.sigma ** 2)))
                        if trace:
                            step.setdefault("obs", {})[o] = {
                                "uv": uv, "box": box, "search": search, "sse": sse, "sse_box": sbox, "sampled": sampled,
                            }
                    terms.append(surface_log_likelihood(model, ps))
                    terms = [x for x in terms if x is not None]
                    if terms:  # otherwise the weights of the last resampling stay (tracker.py:146-149)
                        w = weights_from_log_likelihoods(terms)
                    if resample_method == "stratified":
                        u = random(len(w))
                        idx = stratified_indices(w, u)
                    elif resample_method == "choice":
                        u = random(len(w))
                        idx = choice_indices(w, u)
                    elif resample_method == "residual":
                        idx, u = residual_indices(w, random)
                    else:
                        u = random()
                        idx = systematic_indices(w, u)
                    if trace:
                        step.update({"evolved": ps.copy(), "weights": w.copy(), "u": u, "indices": idx})
                    ps = ps[idx]
                    w = w[idx]
                means[p, t] = weighted_mean(ps, w)
                if return_covariances:
                    sigmas[p, t] = weighted_covariance(ps, w)
                else:
                    sigmas[p, t] = weighted_sigma(ps, w, means[p, t])
                if return_particles:
                    all_ps[p, t], all_w[p, t] = ps, w
                if trace:
                    step.update({"particles": ps.copy(), "post_weights": w.copy()})
                    log.append(step)
        except Exception as exc:  # per-track capture (tracker.py:360-368)
            if raise_errors or P < 2:
                raise
            errors[p] = exc
    return TrackResult(means, sigmas, errors, skipped, all_ps, all_w, log if trace else None, made if trace else None)
