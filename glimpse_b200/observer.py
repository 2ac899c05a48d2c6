"""``Observer``: a time-ordered image sequence from one camera station (reference
``track/observer.py:16-144``).  Animation / plotting helpers are out of scope."""
from __future__ import annotations

import datetime as _dt
from typing import Iterable, List, Union

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

    def cache_images(self, index=slice(None)) -> None:
        for img in np.asarray(self.images, dtype=object)[index]:
            img.read(cache=True)

    def clear_images(self, index=slice(None)) -> None:
        for img in np.asarray(self.images, dtype=object)[index]:
            img.array = None
