"""Frame ingest: one 4288 x 2848 RGB JPEG (quality 92, 4:2:0) decoded by nvJPEG into device memory (gb_decode_jpeg) against
libjpeg-turbo on the host (cv2.imdecode) plus the upload the host path needs.  Prints one JSON line.
    python tools/jpeg_decode_bench.py"""
import json
import os
import sys
import time

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glimpse_b200 import _lib, synthetic  # noqa: E402
from glimpse_b200.image import decode_jpeg  # noqa: E402

torch = _lib.require_cuda()
rng = np.random.RandomState(2)
tex = synthetic.smooth_texture((2848, 4288), rng)
frame = np.ascontiguousarray(np.stack([tex, np.roll(tex, 5, axis=1), np.roll(tex, -3, axis=0)], axis=2))
ok, buf = cv2.imencode(".jpg", frame[:, :, ::-1], [int(cv2.IMWRITE_JPEG_QUALITY), 92])
data = buf.tobytes()
decode_jpeg(data)
t = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = decode_jpeg(data)  # (synchronises before it returns)
    t.append(1e3 * (time.perf_counter() - t0))
cv2.setNumThreads(1)
h = []
for _ in range(3):
    t0 = time.perf_counter()
    ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    dev = torch.from_numpy(ref).cuda()
    torch.cuda.synchronize()
    h.append(1e3 * (time.perf_counter() - t0))
print(json.dumps({"what": "4288 x 2848 RGB JPEG, quality 92, 4:2:0", "jpeg_bytes": len(data), "nvjpeg_to_device_ms": float(np.median(t)),
                  "libjpeg_turbo_plus_upload_ms": float(np.median(h)), "Mpixel_per_s_device": 4288 * 2848 / float(np.median(t)) / 1e3}))
