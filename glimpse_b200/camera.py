"""Host-side mirror of the reference ``glimpse.Camera`` for the Tracker path.

Same constructor arguments, attributes and method names as reference ``camera.py:74-123`` (20-float
parameter vector ``[xyz, viewdir, imgsz, f, c, k1..k6, p1, p2]``), but ``xyz_to_uv`` / ``uv_to_xyz``
run the sm_100a kernels behind ``gb_project`` / ``gb_unproject``.  Calibration, rasterisation and
DEM rendering (reference ``camera.py:720-1129``) are out of scope.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional, Union

import numpy as np

from . import _lib

Vector = Union[Iterable[float], float]


def _fmt(value, length, default=None, dtype=float):
    """helpers.format_list (reference helpers.py:60-100): scalar -> repeated, short -> padded."""
    if value is None:
        return np.full(length, np.nan if default is None else default, dtype=dtype)
    arr = np.atleast_1d(np.asarray(value, dtype=dtype))
    if arr.size == 1 and length > 1 and default is None:
        arr = np.repeat(arr, length)
    if arr.size < length:
        arr = np.concatenate((arr, np.full(length - arr.size, default, dtype=dtype)))
    return arr[:length]


def rotation_matrix(viewdir) -> np.ndarray:
    """``Camera.R`` (reference camera.py:239-280) with NumPy's sin/cos, as the reference computes it."""
    radians = np.deg2rad(np.asarray(viewdir, dtype=float))
    c, s = np.cos(radians), np.sin(radians)
    return np.array(
        [
            [c[0] * c[2] + s[0] * s[1] * s[2], c[0] * s[1] * s[2] - c[2] * s[0], -c[1] * s[2]],
            [c[2] * s[0] * s[1] - c[0] * s[2], s[0] * s[2] + c[0] * c[2] * s[1], -c[1] * c[2]],
            [c[1] * s[0], c[0] * c[1], s[1]],
        ]
    )


def lower_camera(cam) -> _lib.gb_camera:
    """Any object with the reference Camera's public attributes -> ``gb_camera`` struct."""
    out = _lib.gb_camera()
    R = np.asarray(cam.R, dtype=float) if hasattr(cam, "R") else rotation_matrix(cam.viewdir)
    out.R[:] = R.ravel().tolist()
    out.xyz[:] = np.asarray(cam.xyz, dtype=float).tolist()
    out.f[:] = np.asarray(cam.f, dtype=float).tolist()
    imgsz = np.asarray(cam.imgsz).astype(int)
    cc = imgsz / 2 + np.asarray(cam.c, dtype=float)  # camera.py:1507
    out.cc[:] = cc.tolist()
    out.k[:] = np.asarray(cam.k, dtype=float).tolist()
    out.p[:] = np.asarray(cam.p, dtype=float).tolist()
    out.imgsz[:] = imgsz.tolist()
    corr = getattr(cam, "correction", None)
    if isinstance(corr, dict):
        out.has_corr = 1
        out.corr_c1 = corr["refraction"] - 1  # helpers.py:1790
        out.corr_c2 = 2 * corr["radius"]
    return out


class Camera:
    """Distorted frame camera (reference ``camera.py:17-123``)."""

    def __init__(
        self,
        imgsz: Vector,
        f: Vector = None,
        c: Vector = None,
        sensorsz: Vector = None,
        fmm: Vector = None,
        cmm: Vector = None,
        k: Vector = (0, 0, 0, 0, 0, 0),
        p: Vector = (0, 0),
        xyz: Vector = (0, 0, 0),
        viewdir: Vector = (0, 0, 0),
        correction: Union[bool, dict] = False,
    ) -> None:
        if f is not None and fmm is not None:
            raise ValueError("Focal length provided in both pixels and mm (f, fmm)")
        if c is not None and cmm is not None:
            raise ValueError("Principal point offset provided in both pixels and mm (c, cmm)")
        if imgsz is None:
            raise ValueError("Image size (imgsz) cannot be None")
        self._vector = np.full(20, np.nan, dtype=float)
        self.xyz = xyz
        self.viewdir = viewdir
        self.imgsz = imgsz
        self.sensorsz = None if sensorsz is None else _fmt(sensorsz, 2)
        if fmm is not None:
            if self.sensorsz is None:
                raise ValueError("Sensor size is required")
            f = _fmt(fmm, 2) * self.imgsz / self.sensorsz
        if f is None:
            raise ValueError("Focal length (f or fmm) is missing")
        self.f = f
        if cmm is not None:
            if self.sensorsz is None:
                raise ValueError("Sensor size is required")
            c = _fmt(cmm, 2) * self.imgsz / self.sensorsz
        self.c = (0, 0) if c is None else c
        self.k = k
        self.p = p
        if correction is True:
            correction = {}
        if isinstance(correction, dict):
            correction = {"radius": 6.3781e6, "refraction": 0.13, **correction}
        self.correction = correction

    # ---- the 20-vector and its views (camera.py:127-198) ----
    @property
    def vector(self) -> np.ndarray:
        return self._vector

    @property
    def xyz(self):
        return self._vector[0:3]

    @xyz.setter
    def xyz(self, value):
        self._vector[0:3] = _fmt(value, 3, default=0)

    @property
    def viewdir(self):
        return self._vector[3:6]

    @viewdir.setter
    def viewdir(self, value):
        self._vector[3:6] = _fmt(value, 3, default=0)

    @property
    def imgsz(self):
        return self._vector[6:8].astype(int)

    @imgsz.setter
    def imgsz(self, value):
        as_float = _fmt(value, 2)
        if np.any(as_float != np.floor(as_float)):
            raise ValueError("Image size is not integer")
        self._vector[6:8] = as_float

    @property
    def f(self):
        return self._vector[8:10]

    @f.setter
    def f(self, value):
        self._vector[8:10] = _fmt(value, 2)

    @property
    def c(self):
        return self._vector[10:12]

    @c.setter
    def c(self, value):
        self._vector[10:12] = _fmt(value, 2, default=0)

    @property
    def k(self):
        return self._vector[12:18]

    @k.setter
    def k(self, value):
        self._vector[12:18] = _fmt(value, 6, default=0)

    @property
    def p(self):
        return self._vector[18:20]

    @p.setter
    def p(self, value):
        self._vector[18:20] = _fmt(value, 2, default=0)

    @property
    def R(self) -> np.ndarray:
        return rotation_matrix(self.viewdir)

    # ---- projection (camera.py:591-718) ----
    def xyz_to_uv(self, xyz: np.ndarray) -> np.ndarray:
        """World -> image coordinates (n, 2); NaN behind the camera (``gb_project``)."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        pts = torch.as_tensor(np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)).cuda()
        uv = torch.empty((pts.shape[0], 2), dtype=torch.float64, device=pts.device)
        cam = lower_camera(self)
        _lib.check(lib.gb_project(C.byref(cam), pts.data_ptr(), pts.shape[0], uv.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream))
        return uv.cpu().numpy()

    def uv_to_xyz(self, uv: np.ndarray, directions: bool = True, depth: Vector = 1) -> np.ndarray:
        """Image -> world ray directions or points at ``depth`` (``gb_unproject``)."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        pix = torch.as_tensor(np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)).cuda()
        n = pix.shape[0]
        out = torch.empty((n, 3), dtype=torch.float64, device=pix.device)
        dptr = None
        if not isinstance(depth, (int, float)) or depth != 1:
            dep = torch.as_tensor(np.broadcast_to(np.asarray(depth, dtype=np.float64).ravel(), (n,)).copy()).cuda()
            dptr = dep.data_ptr()
        cam = lower_camera(self)
        _lib.check(lib.gb_unproject(C.byref(cam), pix.data_ptr(), n, int(bool(directions)), dptr, out.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream))
        return out.cpu().numpy()

    def inframe(self, uv: np.ndarray) -> np.ndarray:
        """(camera.py:700-718)."""
        with np.errstate(invalid="ignore"):
            return np.all((uv >= 0) & (uv <= self.imgsz), axis=1)

    def copy(self) -> "Camera":
        cam = Camera(imgsz=self.imgsz, f=self.f.copy(), c=self.c.copy(), k=self.k.copy(), p=self.p.copy(),
                     xyz=self.xyz.copy(), viewdir=self.viewdir.copy(),
                     correction=dict(self.correction) if isinstance(self.correction, dict) else self.correction)
        cam.sensorsz = None if self.sensorsz is None else self.sensorsz.copy()
        return cam
