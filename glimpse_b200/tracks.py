"""``Tracks``: result container of ``Tracker.track`` (reference ``track/tracks.py:20-149``).
Plotting / animation are out of scope."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


class Tracks:
    def __init__(self, datetimes, time_unit, means, sigmas=None, covariances=None, particles=None, weights=None,
                 tracker=None, images=None, params: dict = None, errors=None, warnings=None) -> None:
        self.datetimes = np.asarray(datetimes)
        self.time_unit = time_unit
        self.means = means
        self.sigmas = sigmas
        self.covariances = covariances
        self.particles = particles
        self.weights = weights
        self.tracker = tracker
        self.images = images if images is None else np.asarray(images)
        self.params = params
        self.errors = errors if errors is None else np.asarray(errors, dtype=object)
        if warnings is not None:
            w = np.empty(len(warnings), dtype=object)
            for i, item in enumerate(warnings):
                w[i] = item
            warnings = w
        self.warnings = warnings

    @property
    def xyz(self) -> np.ndarray:
        return self.means[:, :, 0:3]

    @property
    def vxyz(self) -> np.ndarray:
        return self.means[:, :, 3:6]

    @property
    def xyz_sigma(self) -> Optional[np.ndarray]:
        if self.sigmas is not None:
            return self.sigmas[:, :, 0:3]
        if self.covariances is not None:
            return np.sqrt(self.covariances[:, :, (0, 1, 2), (0, 1, 2)])
        return None

    @property
    def vxyz_sigma(self) -> Optional[np.ndarray]:
        if self.sigmas is not None:
            return self.sigmas[:, :, 3:6]
        if self.covariances is not None:
            return np.sqrt(self.covariances[:, :, (3, 4, 5), (3, 4, 5)])
        return None

    @property
    def endpoints(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        valid = ~np.isnan(self.means[:, :, 0])
        first = np.argmax(valid, axis=1)
        last = valid.shape[1] - 1 - np.argmax(valid[:, ::-1], axis=1)
        first_valid = valid[np.arange(len(first)), first]
        return first_valid, first[first_valid], last[first_valid]

    @property
    def success(self) -> Optional[np.ndarray]:
        if self.errors is not None:
            return np.array([error is None for error in self.errors])
        return None

    def reverse(self) -> None:
        """Reverse the time order in place (reference tracks.py:131-149)."""
        self.datetimes = self.datetimes[::-1]
        for name in ("means", "sigmas", "covariances", "particles", "weights"):
            value = getattr(self, name)
            if value is not None:
                setattr(self, name, value[:, ::-1])
        if self.images is not None:
            self.images = self.images[::-1]
