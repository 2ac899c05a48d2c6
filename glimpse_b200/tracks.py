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

    # ------------------------------------------------------------------ merging (reference tracks.py:151-203)
    @staticmethod
    def _combine(means: np.ndarray, sigmas: np.ndarray, axis: int, correlated: bool, ignore_nan: bool):
        """Inverse-variance weighted average of normal distributions along ``axis`` (reference
        ``helpers.sum_normals(weights=sigmas**-2, normalize=True)``, helpers.py:523-610): mean = sum w mu,
        variance = sum (w sigma)^2 for uncorrelated terms, (sum w sigma)^2 for fully correlated ones.
        Missing terms are skipped; the result is NaN where any term (``ignore_nan=False``) or every term is missing."""
        missing = np.isnan(means)
        if np.any(missing != np.isnan(sigmas)):
            raise ValueError("Means and sigmas have missing values at different indices")
        if np.any(sigmas == 0):
            raise ValueError("Sigmas cannot be zero")
        with np.errstate(divide="ignore", invalid="ignore"):
            w = np.where(missing, 0.0, sigmas ** -2.0)
            w = w * (1 / w.sum(axis=axis, keepdims=True))
            mu = np.where(missing, 0.0, w * means).sum(axis=axis)
            ws = np.where(missing, 0.0, w * sigmas)
            var = ws.sum(axis=axis) ** 2 if correlated else (ws ** 2).sum(axis=axis)
        blank = missing.all(axis=axis) if ignore_nan else missing.any(axis=axis)
        mu[blank] = np.nan
        var[blank] = np.nan
        return mu, np.sqrt(var)

    @classmethod
    def from_multiple(cls, runs, ignore_nan: bool = False) -> "Tracks":
        """Merge tracks with identical time steps (for example a forward and a backward run): the inverse-variance
        weighted average of their distributions at every time, assumed uncorrelated (reference tracks.py:151-189)."""
        runs = list(runs)
        if len({tuple(run.datetimes) for run in runs}) != 1:
            raise ValueError("Datetimes are not equal for all runs")
        units = {run.time_unit for run in runs}
        if len(units) != 1:
            raise ValueError(f"Time units are not equal for all runs: {units}")
        means = np.stack([run.means for run in runs], axis=3)
        sigmas = np.stack([run.sigmas for run in runs], axis=3)
        mu, sg = cls._combine(means, sigmas, axis=3, correlated=False, ignore_nan=ignore_nan)
        return cls(datetimes=runs[0].datetimes, time_unit=units.pop(), means=mu, sigmas=sg)

    def average(self, ignore_nan: bool = False) -> Tuple[np.ndarray, np.ndarray]:
        """Time-averaged mean and sigma of every track: inverse-variance weights, time steps assumed fully
        correlated (reference tracks.py:191-203)."""
        return self._combine(self.means, self.sigmas, axis=1, correlated=True, ignore_nan=ignore_nan)
