"""Sharded vs single-GPU results of a small fixed problem (torchrun, any world size): prints where they differ.

    python -m torch.distributed.run --nproc-per-node N tools/shard_check.py [points]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import glimpse_b200 as gb  # noqa: E402
from glimpse_b200 import synthetic  # noqa: E402

rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
points = int(sys.argv[1]) if len(sys.argv) > 1 else 64
scene = synthetic.nadir_scene(seed=77, n_points=points, n_particles=2000, n_frames=8, imgsz=(600, 400))
observers, models = synthetic.build(scene, gb)
sharded = gb.Tracker(observers, seed=777).track(models, tile_size=scene.tile_size)
alone = gb.Tracker(observers, seed=777, distributed=False).track(models, tile_size=scene.tile_size)
bad = np.nonzero(~np.all(np.isclose(sharded.means, alone.means, rtol=0, atol=0, equal_nan=True), axis=(1, 2)))[0]
if rank == 0:
    print("world", dist.get_world_size(), "points", points, "differing points", bad.tolist()[:40], "of", points)
    if len(bad):
        p = int(bad[0])
        t = np.nonzero(~np.all(sharded.means[p] == alone.means[p], axis=1))[0]
        print("point", p, "first differing time", t[:3], "max |d| / sigma", np.nanmax(np.abs(sharded.means[p] - alone.means[p]) / alone.sigmas[p]))
        print("errors", [str(e)[:60] for e in sharded.errors if e is not None][:3])
dist.barrier()
dist.destroy_process_group()
