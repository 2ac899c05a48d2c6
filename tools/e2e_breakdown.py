"""Host-side time breakdown of one Tracker.track call on the bench workload (tuning aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import glimpse_b200 as gb
from glimpse_b200 import synthetic, session as S

world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
if world > 1:  # torchrun: every rank tracks its 1000 points of a world x 1000-point scene
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
P, T = 1000 * world, 100
scene = bench.build_scene(P, T, pinned=True)
observers, models = synthetic.build(scene, gb)
tracker = gb.Tracker(observers, seed=1)
for rep in range(3):
    tracker.clear_device_cache()
    torch.cuda.synchronize()
    marks = [("start", time.perf_counter())]
    orig_init = S.Session.__init__

    def timed_init(self, *a, **k):
        orig_init(self, *a, **k)
        marks.append(("session built (uploads enqueued)", time.perf_counter()))

    S.Session.__init__ = timed_init
    orig_run, orig_fetch = S.Session.run, S.Session.fetch

    def timed_run(self):
        orig_run(self)
        marks.append(("gb_track enqueued", time.perf_counter()))

    def timed_fetch(self, *a):
        torch.cuda.synchronize()
        marks.append(("device done", time.perf_counter()))
        out = orig_fetch(self, *a)
        marks.append(("fetched", time.perf_counter()))
        return out

    S.Session.run, S.Session.fetch = timed_run, timed_fetch
    tracks = tracker.track(models, tile_size=scene.tile_size)
    marks.append(("track returned", time.perf_counter()))
    S.Session.__init__, S.Session.run, S.Session.fetch = orig_init, orig_run, orig_fetch
    if rank == 0:
        print(f"rep {rep}: total {1e3 * (marks[-1][1] - marks[0][1]):.1f} ms")
        for (n0, t0), (n1, t1) in zip(marks[:-1], marks[1:]):
            print(f"    {n1:36s} +{1e3 * (t1 - t0):7.1f} ms")
