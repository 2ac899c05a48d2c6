import sys, zlib, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glimpse_b200 as gb
from glimpse_b200 import synthetic
def crc():
    small = synthetic.nadir_scene(seed=77, n_points=64, n_particles=2000, n_frames=8, imgsz=(600, 400))
    observers, models = synthetic.build(small, gb)
    tracks = gb.Tracker(observers, seed=777).track(models, tile_size=small.tile_size)
    c = zlib.crc32(np.ascontiguousarray(tracks.means).tobytes())
    return "%08x" % zlib.crc32(np.ascontiguousarray(tracks.sigmas).tobytes(), c), tracks
a, ta = crc()
print("fresh", a, "errors", sum(e is not None for e in ta.errors))
import torch
junk = torch.full((1 << 28,), float("nan"), dtype=torch.float64, device="cuda"); del junk
junk = torch.full((1 << 28,), 1e300, dtype=torch.float64, device="cuda"); del junk
b, tb = crc()
print("after junk", b, "same", np.array_equal(ta.means, tb.means, equal_nan=True))
for mp in (8, 2):
    small = synthetic.nadir_scene(seed=77, n_points=64, n_particles=2000, n_frames=8, imgsz=(600, 400))
    observers, models = synthetic.build(small, gb)
    t2 = gb.Tracker(observers, seed=777, max_points=mp).track(models, tile_size=small.tile_size)
    d = np.nonzero(~np.all(t2.means == ta.means, axis=(1, 2)))[0]
    print("max_points", mp, "differing points", d.tolist())
