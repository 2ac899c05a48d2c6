"""Image.project on the device (gb_project_image): one D200-size RGB frame (4288 x 2848 x 3 uint8) warped into a camera turned
by a fraction of a degree (a sequence-stabilisation warp), frame resident in HBM, CUDA events around the launches.
Algorithmic bytes per warp = one read of the source + one write of the target.  Prints one JSON line.

    python tools/project_image_bench.py [repeats]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import glimpse_b200 as gb  # noqa: E402
from glimpse_b200 import _lib, synthetic  # noqa: E402
from glimpse_b200.camera import lower_camera  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
torch = _lib.require_cuda()
lib = _lib.load()
dev = torch.device("cuda", 0)
W, H, bands = 4288, 2848, 3
rng = np.random.RandomState(1)
frame = synthetic.smooth_texture((H, W), rng)
frame = np.ascontiguousarray(np.stack([frame, np.roll(frame, 3, axis=1), np.roll(frame, -2, axis=0)], axis=2))
src_cam = gb.Camera(imgsz=(W, H), f=(3700.0, 3690.0), c=(12.5, -8.25), k=synthetic.FULL_K, p=synthetic.FULL_P, xyz=(0, 0, 100.0),
                    viewdir=(60.0, -25.0, 1.5))
dst_cam = gb.Camera(imgsz=(W, H), f=(3700.0, 3690.0), c=(12.5, -8.25), k=synthetic.FULL_K, p=synthetic.FULL_P, xyz=(0, 0, 100.0),
                    viewdir=(60.2, -24.9, 1.4))
out = {}
for method in ("linear", "nearest"):
    pixels = torch.from_numpy(frame.reshape(-1)).to(dev)
    src = _lib.gb_image()
    src.pixels, src.width, src.height, src.pitch, src.nchan, src.dtype = pixels.data_ptr(), W, H, W * bands, bands, _lib.GB_PIX["uint8"]
    src.cam = lower_camera(src_cam)
    dst = lower_camera(dst_cam)
    target = torch.empty(H * W * bands, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # larger than L2: written between the timed launches
    m = 1 if method == "linear" else 0
    for _ in range(3):
        _lib.check(lib.gb_project_image(C.byref(src), C.byref(dst), m, target.data_ptr(), stream))
    ms = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.gb_project_image(C.byref(src), C.byref(dst), m, target.data_ptr(), stream))
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    # end to end through Image.project: host frame in, host frame out
    img = gb.Image("frame", cam=src_cam, datetime=synthetic.T0)
    img.array = frame
    img.project(dst_cam, method=method)
    t0 = time.perf_counter()
    warped = img.project(dst_cam, method=method)
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    nbytes = 2 * W * H * bands
    med = float(np.median(ms))
    out[method] = {"kernel_ms": med, "GB_per_s": nbytes / med / 1e6, "Mpixel_per_s": W * H / med / 1e3, "e2e_ms": e2e_ms,
                   "seen_fraction": float((warped.max(axis=2) > 0).mean())}
peak = 6551.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
print(json.dumps({"what": "Image.project of one 4288 x 2848 x 3 uint8 frame (full k1-k6, p1-p2 distortion in both cameras)",
                  "algorithmic_bytes": 2 * W * H * bands, "hbm_peak_GB_per_s": peak,
                  "roofline_frac_linear": out["linear"]["GB_per_s"] / peak, "results": out}))
