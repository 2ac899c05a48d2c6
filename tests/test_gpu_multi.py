"""Two ranks on two GPUs of one box (NCCL): points are sharded, frames are uploaded once per box and broadcast,
result blocks are gathered on the devices — and every rank gets exactly what a single GPU computes."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
    import glimpse_b200 as gb
    from glimpse_b200 import synthetic
    rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    scene = synthetic.nadir_scene(seed=4, n_points=7, n_particles=1500, n_frames=6, imgsz=(400, 300), margin_px=100)
    observers, models = synthetic.build(scene, gb)
    sharded = gb.Tracker(observers, seed=11).track(models, tile_size=scene.tile_size, return_covariances=True)
    h2d_sharded = sharded.tracker.last_run["h2d_bytes"]
    alone = gb.Tracker(observers, seed=11, distributed=False).track(models, tile_size=scene.tile_size, return_covariances=True)
    assert all(e is None for e in sharded.errors)
    np.testing.assert_array_equal(sharded.means, alone.means)          # Philox counters use global point indices
    np.testing.assert_array_equal(sharded.covariances, alone.covariances)
    assert h2d_sharded < 0.7 * alone.tracker.last_run["h2d_bytes"]     # half of the frames came over NVLink
    # blocks of points inside every rank (a track too large for the device memory): same answer again
    # (uneven: rank 0 holds 4 points = two sessions of <= 3, rank 1 holds 3 = one session; no collective depends on it)
    blocked = gb.Tracker(observers, seed=11, max_points=3)
    tracks = blocked.track(models, tile_size=scene.tile_size, return_covariances=True)
    assert blocked.last_run.get("sessions", 1) == (2 if rank == 0 else 1)
    np.testing.assert_array_equal(tracks.means, alone.means)
    np.testing.assert_array_equal(tracks.covariances, alone.covariances)
    # seed=None: one key for the whole call, rank 0's draw on every rank; particles returned through the packed gather
    np.random.seed(100 + rank)
    free = gb.Tracker(observers).track(models, tile_size=scene.tile_size, return_particles=True)
    ref = gb.Tracker(observers, distributed=False, seed=None)
    np.random.seed(100)
    alone2 = ref.track(models, tile_size=scene.tile_size, return_particles=True)
    np.testing.assert_array_equal(free.means, alone2.means)
    np.testing.assert_array_equal(free.particles, alone2.particles)
    dist.barrier(); dist.destroy_process_group()
    print("rank", rank, "ok")
''')


def test_two_ranks_match_one_gpu(cuda, tmp_path):
    if cuda.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), ROOT]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("ok") == 2
