"""glimpse_b200 — B200-native (sm_100a) implementation of the ``glimpse.Tracker`` hot path.

Public names mirror the reference package (``glimpse.Camera``, ``Image``, ``Raster``, ``Observer``,
``CartesianMotion``, ``CylindricalMotion``, ``TangentCartesianMotion``, ``TangentCylindricalMotion``, ``Tracker``, ``Tracks``).  All compute goes through the
C ABI in ``include/glimpse_b200.h`` (``libglimpse_b200.so``); there is no CPU fallback.
"""
from .camera import Camera
from .image import Image, Raster
from .motion import CartesianMotion, CylindricalMotion, TangentCartesianMotion, TangentCylindricalMotion
from .observer import Observer
from .tracker import Tracker
from .tracks import Tracks

__all__ = ["Camera", "Image", "Raster", "Observer", "CartesianMotion", "CylindricalMotion", "TangentCartesianMotion",
           "TangentCylindricalMotion", "Tracker", "Tracks"]
__version__ = "0.1.0"
