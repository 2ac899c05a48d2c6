"""cProfile of Session construction on the bench workload (tuning aid)."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import glimpse_b200 as gb
from glimpse_b200 import synthetic

scene = bench.build_scene(1000, 100, pinned=True)
observers, models = synthetic.build(scene, gb)
tracker = gb.Tracker(observers, seed=1)
for _ in range(2):
    tracker.clear_device_cache()
    tracker.track(models, tile_size=scene.tile_size)
tracker.clear_device_cache()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
tracker.track(models, tile_size=scene.tile_size)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
