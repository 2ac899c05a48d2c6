"""Import shim for the *real* reference (ezwelty/glimpse) in the build container.

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` (and nothing
else) to import the unmodified reference from ``/root/reference/src`` so that
golden vectors can be generated from the reference itself.  The reference is
pure Python but imports six third-party modules that are not installed here
(GDAL, sharedmem, matplotlib, lmfit, piexif, progress); none of them is executed
on the ``Tracker.track`` path when frames are supplied in memory, so they are
replaced by inert stand-ins (recipe: SURVEY.md Appendix B).

``/root/reference`` does not exist on the GPU box: nothing under ``tests/``
marked ``gpu``, ``bench.py`` or ``__graft_entry__.py`` imports this module.
"""
import os
import sys
import types

import numpy as np

REFERENCE_SRC = os.environ.get("GLIMPSE_REFERENCE_SRC", "/root/reference/src")


class _Inert:
    """Object that absorbs attribute access and calls made at import time."""

    def __getattr__(self, key):
        return _Inert()

    def __call__(self, *args, **kwargs):
        return _Inert()


class _SerialPool:
    """Serial stand-in for ``sharedmem.MapReduce`` (its np=0 behaviour)."""

    def __init__(self, np=0):
        self.np = np

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def map(self, func, sequence, reduce=None, star=False):
        results = []
        for item in sequence:
            value = func(*item) if star else func(item)
            if reduce is not None:
                value = reduce(*value) if isinstance(value, tuple) else reduce(value)
            results.append(value)
        return results


_MISSING = [
    "osgeo", "osgeo.gdal", "osgeo.gdal_array", "osgeo.ogr", "osgeo.osr",
    "matplotlib", "matplotlib.animation", "matplotlib.patches", "matplotlib.pyplot",
    "matplotlib.axes", "matplotlib.quiver", "matplotlib.colors", "matplotlib.image",
    "matplotlib.lines", "matplotlib.container", "matplotlib.collections", "matplotlib.path",
    "lmfit", "piexif", "progress", "progress.bar",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "glimpse"))


def import_reference():
    """Return the reference ``glimpse`` package (raises if it is not on this machine)."""
    if "glimpse" in sys.modules and getattr(sys.modules["glimpse"], "__file__", "").startswith(REFERENCE_SRC):
        return sys.modules["glimpse"]
    if not available():
        raise ImportError(f"reference sources not found under {REFERENCE_SRC}")
    for name in _MISSING:
        try:
            __import__(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        mod.__path__ = []
        mod.__getattr__ = lambda key: _Inert()
        sys.modules[name] = mod
    for name in list(sys.modules):
        if "." in name and name.split(".")[0] in ("osgeo", "matplotlib", "progress"):
            parent, child = name.rsplit(".", 1)
            if parent in sys.modules:
                setattr(sys.modules[parent], child, sys.modules[name])
    if "sharedmem" not in sys.modules:
        sm = types.ModuleType("sharedmem")
        sm.MapReduce = sm.MapReduceByThread = _SerialPool
        sm.copy = lambda a: np.array(a)
        sys.modules["sharedmem"] = sm
    sys.path.insert(0, REFERENCE_SRC)
    import glimpse  # noqa: E402

    return glimpse
