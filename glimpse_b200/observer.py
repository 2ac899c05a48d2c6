"""``Observer``: a time-ordered image sequence from one camera station (reference
``track/observer.py:16-214, 455-493``).  Animation / plotting helpers are out of scope."""
from __future__ import annotations

import datetime as _dt
from typing import Any, Iterable, List, Union

import numpy as np


class Observer:
    def __init__(self, images: Iterable, sigma: float = 0.3, cache: bool = True) -> None:
        images = list(images)
        if len(images) < 2:
            raise ValueError("Images are not two or greater")
        datetimes = []
        for i, img in enumerate(images):
            if img.datetime is None:
                raise ValueError(f"Image {i} is missing datetime")
            datetimes.append(img.datetime)
        deltas = np.array([dt.total_seconds() for dt in np.diff(datetimes)])
        if any(deltas <= 0):
            raise ValueError("Image datetimes are not stricly increasing")
        self.images: List = images
        self.datetimes = np.array(datetimes)
        self.sigma = sigma
        self.cache = cache

    def index(self, value, maxdt: _dt.timedelta = _dt.timedelta(0)) -> int:
        """(reference observer.py:71-100)."""
        if isinstance(value, _dt.datetime):
            dts = np.abs(value - self.datetimes)
            index = int(np.argmin(dts))
            if maxdt is not None and dts[index] > abs(maxdt):
                raise ValueError("Nearest image out of range by " + str(dts[index] - abs(maxdt)))
            return index
        return self.images.index(value)

    def xyz_to_uv(self, xyz: np.ndarray, img: int) -> np.ndarray:
        return self.images[img].xyz_to_uv(xyz)

    def tile_box(self, uv: Iterable[float], size: Iterable[int], img: int) -> np.ndarray:
        """Grid-aligned box around ``uv`` (reference observer.py:115-130, raster.py:390-421)."""
        half = np.multiply(size, 0.5)
        corners = np.vstack((np.asarray(uv, dtype=float) - half, np.asarray(uv, dtype=float) + half))
        if not np.all((corners >= 0) & (corners <= np.asarray(self.images[img].size))):
            raise IndexError("Box extends beyond grid bounds")
        return np.floor(corners + 0.5).flatten().astype(int)

    def extract_tile(self, box: Iterable[int], img: int) -> np.ndarray:
        return self.images[img].read(box=box, cache=self.cache)

    def cache_images(self, index=slice(None), device: bool = False) -> None:
        """Read the images ahead of tracking (reference observer.py:260-268).  ``device=True``: JPEG files are decoded by
        nvJPEG straight into device memory (``Image.read_device``) and tracked from there without an upload."""
        for img in np.asarray(self.images, dtype=object)[index]:
            if device and hasattr(img, "read_device"):
                img.read_device(cache=True)
            else:
                img.read(cache=True)

    def clear_images(self, index=slice(None)) -> None:
        for img in np.asarray(self.images, dtype=object)[index]:
            img.array = None
            if hasattr(img, "device_array"):
                img.device_array = None

    # ------------------------------------------------------------------ sub-pixel sampling (device spline)
    @staticmethod
    def _spline_degrees(kwargs: dict):
        kx, ky = kwargs.get("kx", 3), kwargs.get("ky", 3)
        if set(kwargs) - {"kx", "ky"} or kx not in (1, 2, 3, 4, 5) or ky not in (1, 2, 3, 4, 5):
            raise NotImplementedError("only RectBivariateSpline(kx = 1..5, ky = 1..5, s = 0) has a device kernel")
        return int(kx), int(ky)

    @staticmethod
    def _sample(tile: np.ndarray, rows: np.ndarray, cols: np.ndarray, kx: int, ky: int) -> np.ndarray:
        """Spline through the cell centres of a 2-D ``tile`` at (row, column) coordinates in cell units from the first centre
        (``gb_sample_surface``: the same Hermite-form spline the tracker samples its SSE surfaces with)."""
        from . import _lib

        torch = _lib.require_cuda()
        lib = _lib.load()
        dev = torch.device("cuda", torch.cuda.current_device())
        t = torch.as_tensor(np.ascontiguousarray(tile, dtype=float)).to(dev)
        xy = torch.as_tensor(np.ascontiguousarray(np.column_stack((cols, rows)), dtype=float)).to(dev)
        work = torch.empty((tile.shape[0] * (tile.shape[1] | 1) * 16 + (tile.shape[0] + tile.shape[1]) * 88,), dtype=torch.uint8, device=dev)
        out = torch.empty((xy.shape[0],), dtype=torch.float64, device=dev)
        _lib.check(lib.gb_sample_surface(t.data_ptr(), tile.shape[0], tile.shape[1], kx, ky, xy.data_ptr(), xy.shape[0], work.data_ptr(),
                                         out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return out.cpu().numpy()

    def sample_tile(self, uv, tile: np.ndarray, box: Iterable[float], grid: bool = False, **kwargs: Any) -> np.ndarray:
        """Sample ``tile`` (ny, nx), whose cells span ``box`` (left, top, right, bottom), at image coordinates: points
        (n, [u, v]), or grid vectors [(nu,), (nv,)] with ``grid=True`` (reference observer.py:178-214)."""
        kx, ky = self._spline_degrees(kwargs)
        box = np.asarray(box, dtype=float)
        if grid:
            uu, vv = np.asarray(uv[0], dtype=float), np.asarray(uv[1], dtype=float)
            pts = np.column_stack((np.tile(uu, len(vv)), np.repeat(vv, len(uu))))
        else:
            pts = np.atleast_2d(np.asarray(uv, dtype=float))
        inside = (pts[:, 0] >= box[0]) & (pts[:, 0] <= box[2]) & (pts[:, 1] >= box[1]) & (pts[:, 1] <= box[3])
        if not np.all(inside):
            raise ValueError("Some sampling points are outside box")
        du = (box[2] - box[0]) / tile.shape[1]
        dv = (box[3] - box[1]) / tile.shape[0]
        out = self._sample(tile, (pts[:, 1] - (box[1] + dv * 0.5)) / dv, (pts[:, 0] - (box[0] + du * 0.5)) / du, kx, ky)
        return out.reshape(len(uv[1]), len(uv[0])) if grid else out

    def shift_tile(self, tile: np.ndarray, duv: Iterable[float], **kwargs: Any) -> np.ndarray:
        """Shift ``tile`` (2-D or 3-D) by a sub-pixel offset (du, dv) of at most half a pixel (reference observer.py:146-176)."""
        if any(np.abs(duv) > 0.5):
            raise ValueError("Shift larger than 0.5 pixels")
        kx, ky = self._spline_degrees(kwargs)
        bands = np.atleast_3d(tile).astype(float)
        ny, nx = bands.shape[:2]
        rows = np.repeat(np.arange(ny, dtype=float) + duv[1], nx)
        cols = np.tile(np.arange(nx, dtype=float) + duv[0], ny)
        for i in range(bands.shape[2]):
            bands[:, :, i] = self._sample(bands[:, :, i], rows, cols, kx, ky).reshape(ny, nx)
        return bands.squeeze(axis=2) if bands.shape[2] == 1 else bands

    # ------------------------------------------------------------------ sub-sequences
    def subset(self, **kwargs: Any) -> "Observer":
        """New Observer with the images selected by ``select_datetimes`` (reference observer.py:455-464)."""
        mask = select_datetimes(self.datetimes, **kwargs)
        images = [img for img, keep in zip(self.images, mask) if keep]
        return self.__class__(images, sigma=self.sigma, cache=self.cache)

    def split(self, n, overlap: int = 1) -> List["Observer"]:
        """Split into ``n`` equal-length Observers or at datetime breaks (reference observer.py:466-493)."""
        if np.iterable(n):
            breaks = np.unique(np.hstack((n, self.datetimes[[0, -1]])))
        else:
            dt = (self.datetimes[-1] - self.datetimes[0]) / n
            breaks = datetime_range(self.datetimes[0], self.datetimes[-1], dt)
        observers = []
        start = breaks[0]
        for i in range(len(breaks) - 1):
            observer = self.subset(start=start, end=breaks[i + 1])
            if overlap:
                lag = min(overlap, len(observer.datetimes))
                start = observer.datetimes[-lag]
            else:
                start = observer.datetimes[-1] + _dt.timedelta(microseconds=1)
            observers.append(observer)
        return observers


def datetime_range(start: _dt.datetime, stop: _dt.datetime, step: _dt.timedelta) -> np.ndarray:
    """Datetimes from ``start`` to ``stop`` inclusive (reference helpers.py:1856-1880)."""
    max_steps = (stop - start) // step
    return np.array([start + n * step for n in range(max_steps + 1)])


def select_datetimes(datetimes, start=None, end=None, snap=None, maxdt=None, origin=_dt.datetime(1970, 1, 1)) -> np.ndarray:
    """Boolean mask of the datetimes in [start, end], optionally only those nearest to the multiples of ``snap`` (within
    ``maxdt``, default half of ``snap``) counted from ``origin`` (reference helpers.py:1883-1958)."""
    datetimes = np.asarray(datetimes)
    selected = np.ones(datetimes.shape, dtype=bool)
    if start:
        selected &= datetimes >= start
    else:
        start = datetimes[0]
        if snap:
            start -= snap
    if end:
        selected &= datetimes <= end
    else:
        end = datetimes[-1]
        if snap:
            end += snap
    if start > end:
        raise ValueError("Start datetime is after end datetime")
    if snap:
        shift = (origin - start) % snap
        start = start + shift
        targets = datetime_range(start, end, step=snap)
        nearest = find_nearest_datetimes(targets, datetimes)
        if maxdt is None:
            maxdt = snap * 0.5
        distances = np.abs(targets - datetimes[nearest])
        nearest = np.unique(nearest[distances <= maxdt])
        temp = np.zeros(datetimes.shape, dtype=bool)
        temp[nearest] = True
        selected &= temp
    return selected


def find_nearest_datetimes(a, b) -> np.ndarray:
    """For each datetime of ``a`` the index of the nearest datetime of ``b``."""
    at = np.array([v.timestamp() for v in a], dtype=float)
    bt = np.array([v.timestamp() for v in b], dtype=float)
    return np.argmin(np.abs(at[:, None] - bt[None, :]), axis=1)
