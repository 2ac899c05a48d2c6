"""cProfile of one warm end-to-end Tracker.track of the bench workload (host view): where the milliseconds outside the
kernels go.    python tools/e2e_profile.py [points] [frames]"""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import glimpse_b200 as gb  # noqa: E402
from glimpse_b200 import synthetic  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
scene = bench.build_scene(P, T, pinned=True)
observers, models = synthetic.build(scene, gb)
tracker = gb.Tracker(observers, seed=1)
for _ in range(3):
    tracker.clear_device_cache()
    tracker.track(models, tile_size=scene.tile_size)
print("host_ms", {k: round(v, 2) for k, v in tracker.last_run["host_ms"].items()})
prof = cProfile.Profile()
tracker.clear_device_cache()
prof.enable()
tracker.track(models, tile_size=scene.tile_size)
prof.disable()
pstats.Stats(prof).sort_stats("cumulative").print_stats(28)
