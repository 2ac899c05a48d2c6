"""Where the host time of Session.__init__ goes for the bench workload (1 000 models, 100 frames): wall-clock laps around its
parts, three warm calls.    python tools/session_laps.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import glimpse_b200 as gb
from glimpse_b200 import synthetic, session as S
scene = bench.build_scene(1000, 100, pinned=True)
observers, models = synthetic.build(scene, gb)
tracker = gb.Tracker(observers, seed=1)
for _ in range(3):
    tracker.clear_device_cache(); tracker.track(models, tile_size=scene.tile_size)
# time pieces of Session.__init__ by wrapping methods
import types
laps = {}
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); laps[name] = laps.get(name, 0) + 1e3*(time.perf_counter()-t0); return r
    setattr(obj, name, g)
for name in ("_upload_frames", "_start_frame_copies", "_lower_models"):
    wrap(S.Session, name)
wrap(S, "lower_models"); wrap(S, "point_span"); wrap(S, "frames_need_ranks")
orig_init = S.Session.__init__
def init(self, *a, **k):
    t0 = time.perf_counter(); orig_init(self, *a, **k); laps["__init__"] = 1e3*(time.perf_counter()-t0)
S.Session.__init__ = init
for _ in range(3):
    laps.clear()
    tracker.clear_device_cache(); tracker.track(models, tile_size=scene.tile_size)
    print({k: round(v, 2) for k, v in laps.items()}, {k: round(v,2) for k,v in tracker.last_run["host_ms"].items()})
