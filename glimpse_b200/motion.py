"""Motion models (reference ``track/motion.py:92-522``) as parameter holders that the Tracker lowers
to ``gb_motion`` structs.  The Motion protocol of the reference (motion.py:13-89: ``initialize_particles``,
``evolve_particles``, ``compute_log_likelihoods``) is also callable on its own, each method through its own entry point
(``gb_init_particles``, ``gb_evolve``, ``gb_motion_log_likelihoods``) with draws from the legacy global NumPy generator in
the reference's order."""
from __future__ import annotations

import ctypes as C
import datetime as _dt
from typing import Iterable, Union

import numpy as np

from . import _lib
from .image import Raster

Number = Union[int, float]


def _as_raster(value) -> Raster:
    if hasattr(value, "array") and hasattr(value, "xlim"):
        return value
    if value is None:
        # reference motion.py:139-141 wraps None as well, and then fails on first use
        raise ValueError("dem_sigma=None is not usable (the reference raises KeyError on it); pass 0")
    return Raster(value, x=(-np.inf, np.inf), y=(-np.inf, np.inf))


class CartesianMotion:
    """Random-acceleration model with independent x, y, z components (reference motion.py:92-204)."""

    kind = _lib.GB_MOTION_CARTESIAN

    def __init__(self, xy, time_unit: _dt.timedelta, dem, dem_sigma=0.0, n: int = 1000, xy_sigma=(0, 0),
                 vxyz=(0, 0, 0), vxyz_sigma=(0, 0, 0), axyz=(0, 0, 0), axyz_sigma=(0, 0, 0)) -> None:
        self.xy = xy
        self.time_unit = time_unit
        self.dem = _as_raster(dem)
        self.dem_sigma = _as_raster(dem_sigma)
        self.n = n
        self.xy_sigma = xy_sigma
        self.vxyz = vxyz
        self.vxyz_sigma = vxyz_sigma
        self.axyz = axyz
        self.axyz_sigma = axyz_sigma

    def _velocity(self):
        return self.vxyz, self.vxyz_sigma, self.axyz, self.axyz_sigma

    def lower(self, dem_index: int, dem_sigma_index: int) -> _lib.gb_motion:
        m = _lib.gb_motion()
        m.kind = self.kind
        m.dem, m.dem_sigma = dem_index, dem_sigma_index
        v, vs, a, as_ = self._velocity()
        m.xy[:] = np.asarray(self.xy, dtype=float).tolist()
        m.xy_sigma[:] = np.broadcast_to(np.asarray(self.xy_sigma, dtype=float), (2,)).tolist()
        m.v[:] = np.asarray(v, dtype=float).tolist()
        m.v_sigma[:] = np.asarray(vs, dtype=float).tolist()
        m.a[:] = np.asarray(a, dtype=float).tolist()
        m.a_sigma[:] = np.asarray(as_, dtype=float).tolist()
        m.slope_sigma = float(getattr(self, "slope_sigma", 0.0))
        return m

    def _draw_step_normals(self, n: int) -> np.ndarray:
        """The reference's draws of one ``evolve_particles`` call, as (n, 3)."""
        return np.random.randn(n, 3)

    def _draw_init_normals(self, n: int) -> np.ndarray:
        """The reference's draws of one ``initialize_particles`` call, as (n, 6): randn(n, 2), randn(n), randn(n, 3)."""
        z = np.empty((n, 6))
        z[:, 0:2] = np.random.randn(n, 2)
        z[:, 2] = np.random.randn(n)
        z[:, 3:6] = np.random.randn(n, 3)
        return z

    def _device_tables(self, torch):
        """(motion struct, surface table, keep-alive tensors) of this one model on the current device."""
        device = torch.device("cuda", torch.cuda.current_device())
        s_dem, t_dem = self.dem.lower(torch, device)
        s_sig, t_sig = self.dem_sigma.lower(torch, device)
        surfaces = torch.frombuffer(bytearray(bytes(s_dem) + bytes(s_sig)), dtype=torch.uint8).cuda()
        motion = torch.frombuffer(bytearray(bytes(self.lower(0, 1))), dtype=torch.uint8).cuda()
        return motion, surfaces, (t_dem, t_sig)

    @staticmethod
    def _raise_status(status) -> None:
        code = int(status.item())
        if code:
            cls, msg = _lib.GB_ST_MESSAGES[code]  # DEM sampled out of bounds: raster.py:961-973
            raise cls(msg)

    def initialize_particles(self) -> np.ndarray:
        """(n, 6) particles around the initial position and velocity (motion.py:149-163, 260-283, 378-390, 485-505),
        evaluated by ``gb_init_particles``."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        n = int(self.n)
        normals = torch.as_tensor(np.ascontiguousarray(self._draw_init_normals(n))).cuda()
        state = torch.empty((6, n), dtype=torch.float64, device=normals.device)
        motion, surfaces, keep = self._device_tables(torch)
        status = torch.zeros(1, dtype=torch.int32, device=normals.device)
        _lib.check(lib.gb_init_particles(motion.data_ptr(), surfaces.data_ptr(), 1, n, normals.data_ptr(), state.data_ptr(),
                                         status.data_ptr(), torch.cuda.current_stream().cuda_stream))
        self._raise_status(status)
        del keep
        return np.ascontiguousarray(state.cpu().numpy().T)

    def compute_log_likelihoods(self, particles: np.ndarray):
        """(n,) log-likelihoods of the particles' heights above the DEM (motion.py:181-204), evaluated by
        ``gb_motion_log_likelihoods``; the tangent models have none (motion.py:77-89) and return ``None``."""
        if self.kind in (_lib.GB_MOTION_TANGENT_CARTESIAN, _lib.GB_MOTION_TANGENT_CYLINDRICAL):
            return None
        torch = _lib.require_cuda()
        lib = _lib.load()
        n = len(particles)
        state = torch.as_tensor(np.ascontiguousarray(np.asarray(particles, dtype=float).T)).cuda()  # [6][n]
        ll = torch.empty(n, dtype=torch.float64, device=state.device)
        motion, surfaces, keep = self._device_tables(torch)
        status = torch.zeros(1, dtype=torch.int32, device=state.device)
        _lib.check(lib.gb_motion_log_likelihoods(motion.data_ptr(), surfaces.data_ptr(), 1, n, state.data_ptr(), ll.data_ptr(),
                                                 status.data_ptr(), torch.cuda.current_stream().cuda_stream))
        self._raise_status(status)
        del keep
        return ll.cpu().numpy()

    def evolve_particles(self, particles: np.ndarray, dt: _dt.timedelta) -> None:
        """In-place motion step on (n, 6) particles with draws from the legacy global NumPy generator in the
        reference's order (motion.py:165-179, 285-311, 392-420, 507-522), evaluated by ``gb_evolve``."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        n = len(particles)
        tau = dt.total_seconds() / self.time_unit.total_seconds()
        normals = torch.as_tensor(np.ascontiguousarray(self._draw_step_normals(n))).cuda()
        state = torch.as_tensor(np.ascontiguousarray(particles.T)).cuda()  # [6][n]
        motion, surfaces, keep = self._device_tables(torch)
        status = torch.zeros(1, dtype=torch.int32, device=state.device)
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.gb_evolve(motion.data_ptr(), surfaces.data_ptr(), 1, n, tau, tau ** 2, normals.data_ptr(),
                                 state.data_ptr(), status.data_ptr(), stream))
        self._raise_status(status)
        particles[:] = state.cpu().numpy().T
        del keep


class CylindricalMotion(CartesianMotion):
    """Speed / heading / elevation components (reference motion.py:207-311)."""

    kind = _lib.GB_MOTION_CYLINDRICAL

    def __init__(self, xy, time_unit: _dt.timedelta, dem, dem_sigma=0.0, n: int = 1000, xy_sigma=(0, 0),
                 vrthz=(0, 0, 0), vrthz_sigma=(0, 0, 0), arthz=(0, 0, 0), arthz_sigma=(0, 0, 0)) -> None:
        self.xy = xy
        self.time_unit = time_unit
        self.dem = _as_raster(dem)
        self.dem_sigma = _as_raster(dem_sigma)
        self.n = n
        self.xy_sigma = xy_sigma
        self.vrthz = vrthz
        self.vrthz_sigma = vrthz_sigma
        self.arthz = arthz
        self.arthz_sigma = arthz_sigma

    def _velocity(self):
        return self.vrthz, self.vrthz_sigma, self.arthz, self.arthz_sigma


class TangentCartesianMotion(CartesianMotion):
    """Particles move tangent to a mean surface: horizontal random-acceleration model, heights follow the DEM plus
    a per-particle offset that random-walks with the distance travelled (reference motion.py:314-420)."""

    kind = _lib.GB_MOTION_TANGENT_CARTESIAN

    def __init__(self, xy, time_unit: _dt.timedelta, dem, dem_sigma=0.0, n: int = 1000, xy_sigma=(0, 0),
                 vxy=(0, 0), vxy_sigma=(0, 0), axy=(0, 0), axy_sigma=(0, 0), slope_sigma: Number = 0) -> None:
        self.xy = xy
        self.time_unit = time_unit
        self.dem = _as_raster(dem)
        self.dem_sigma = _as_raster(dem_sigma)
        self.n = n
        self.xy_sigma = xy_sigma
        self.vxy = vxy
        self.vxy_sigma = vxy_sigma
        self.axy = axy
        self.axy_sigma = axy_sigma
        self.slope_sigma = slope_sigma

    def _velocity(self):
        pad = lambda x: tuple(np.broadcast_to(np.asarray(x, dtype=float), (2,))) + (0.0,)  # noqa: E731
        return pad(self.vxy), pad(self.vxy_sigma), pad(self.axy), pad(self.axy_sigma)

    def _draw_init_normals(self, n: int) -> np.ndarray:
        """randn(n, 2), randn(n), randn(n, 2): the tangent models have no vertical velocity (motion.py:378-390, 485-505)."""
        z = np.zeros((n, 6))
        z[:, 0:2] = np.random.randn(n, 2)
        z[:, 2] = np.random.randn(n)
        z[:, 3:5] = np.random.randn(n, 2)
        return z

    def _draw_step_normals(self, n: int) -> np.ndarray:
        first = np.random.randn(n, 2)  # accelerations
        return np.column_stack((first, np.random.randn(n)))  # then the slope walk (motion.py:400-407)


class TangentCylindricalMotion(TangentCartesianMotion):
    """``TangentCartesianMotion`` with speed / heading components (reference motion.py:423-522)."""

    kind = _lib.GB_MOTION_TANGENT_CYLINDRICAL

    def __init__(self, xy, time_unit: _dt.timedelta, dem, dem_sigma=0.0, n: int = 1000, xy_sigma=(0, 0),
                 vrth=(0, 0), vrth_sigma=(0, 0), arth=(0, 0), arth_sigma=(0, 0), slope_sigma: Number = 0) -> None:
        self.xy = xy
        self.time_unit = time_unit
        self.dem = _as_raster(dem)
        self.dem_sigma = _as_raster(dem_sigma)
        self.n = n
        self.xy_sigma = xy_sigma
        self.vrth = vrth
        self.vrth_sigma = vrth_sigma
        self.arth = arth
        self.arth_sigma = arth_sigma
        self.slope_sigma = slope_sigma

    def _velocity(self):
        pad = lambda x: tuple(np.broadcast_to(np.asarray(x, dtype=float), (2,))) + (0.0,)  # noqa: E731
        return pad(self.vrth), pad(self.vrth_sigma), pad(self.arth), pad(self.arth_sigma)
