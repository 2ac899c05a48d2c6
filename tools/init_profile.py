import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench, glimpse_b200 as gb
from glimpse_b200 import synthetic, session as S
scene = bench.build_scene(1000, 100, pinned=True)
observers, models = synthetic.build(scene, gb)
tracker = gb.Tracker(observers, seed=1)
acc = {}
def wrap(name):
    orig = getattr(S.Session, name)
    def f(self, *a, **k):
        t0 = time.perf_counter(); r = orig(self, *a, **k); acc[name] = acc.get(name, 0) + time.perf_counter() - t0; return r
    setattr(S.Session, name, f)
for n in ("_lower_models", "_upload_frames", "_start_frame_copies", "__init__", "run", "fetch", "final_state"):
    wrap(n)
for rep in range(10):
    tracker.clear_device_cache(); torch.cuda.synchronize(); acc.clear()
    t0 = time.perf_counter(); tracker.track(models, tile_size=scene.tile_size); tot = time.perf_counter() - t0
    print(rep, f"total {tot*1e3:.1f}", {k: round(v*1e3, 2) for k, v in acc.items()})
