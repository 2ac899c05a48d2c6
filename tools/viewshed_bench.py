"""Raster.viewshed on the device (gb_viewshed): a seeded 2000 x 2000 DEM seen from a point near its middle, curvature and
refraction corrected.  Prints one JSON line: end-to-end time of Raster.viewshed (host array in, boolean array out), the device
time of the seven kernels alone (CUDA events, DEM resident), visible fraction.   python tools/viewshed_bench.py [cells per side]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import glimpse_b200 as gb  # noqa: E402
from glimpse_b200 import _lib  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
torch = _lib.require_cuda()
lib = _lib.load()
dev = torch.device("cuda", 0)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402

case = scenes.viewshed_large_case(side)
z, origin = case["z"], case["origin"]
raster = gb.Raster(z, x=case["xlim"], y=case["ylim"])
vis = raster.viewshed(origin, correction=True)  # warm-up (module load, allocator)
t0 = time.perf_counter()
vis = raster.viewshed(origin, correction=True)
e2e_ms = 1e3 * (time.perf_counter() - t0)

# device time alone: the same call on resident arrays
x, y = raster.x, raster.y
max_rings = int(np.hypot(side, side)) + 8
nbytes = int(lib.gb_viewshed_work_bytes(side, side, max_rings))
work = torch.empty(nbytes, dtype=torch.uint8, device=dev)
z_d, x_d, y_d = (torch.from_numpy(np.array(v, dtype=float, order="C", copy=True)).to(dev) for v in (z, x, y))
out = torch.empty(side * side, dtype=torch.uint8, device=dev)
corr = (C.c_double * 2)(6.3781e6, 0.13)
ms = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(lib.gb_viewshed(z_d.data_ptr(), side, side, x_d.data_ptr(), y_d.data_ptr(), 10.0, (C.c_double * 3)(*origin), corr,
                               max_rings, work.data_ptr(), nbytes, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    b.record()
    torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
assert int(work[:4].cpu().numpy().view(np.int32)[0]) == 0
assert np.array_equal(out.cpu().numpy().reshape(side, side).astype(bool), vis)
print(json.dumps({"what": f"Raster.viewshed of a {side} x {side} DEM (curvature / refraction corrected), origin near the middle",
                  "cells": side * side, "device_ms": float(np.median(ms)), "e2e_ms": e2e_ms, "Mcells_per_s_device": side * side / float(np.median(ms)) / 1e3,
                  "visible_fraction": float(vis.mean())}))
