"""Per-call, per-rank view of repeated end-to-end Tracker.track calls of the bench workload (under torchrun or alone):
wall time, the host laps of Tracker.last_run["host_ms"], and what the CUDA caching allocators did during the call
(cudaMalloc / cudaHostAlloc counts), first with the cyclic collector disabled (as bench.py times), then enabled.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/e2e_steps.py [calls]"""
import gc
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import glimpse_b200 as gb  # noqa: E402
from glimpse_b200 import synthetic  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 8
P, T = bench.WORKLOAD["n_points"], bench.WORKLOAD["n_frames"]
scene = bench.build_scene(P * world, T, pinned=True)
observers, models = synthetic.build(scene, gb)
tracker = gb.Tracker(observers, seed=20260101)


def counters():
    dev = torch.cuda.memory_stats()
    host = torch.cuda.host_memory_stats() if hasattr(torch.cuda, "host_memory_stats") else {}
    return {"dev_alloc": dev.get("num_device_alloc", 0), "dev_free": dev.get("num_device_free", 0),
            "dev_reserved_mb": dev.get("reserved_bytes.all.current", 0) >> 20,
            "host_alloc": host.get("num_host_alloc", 0), "host_free": host.get("num_host_free", 0),
            "host_mb": host.get("allocated_bytes.current", host.get("reserved_bytes.current", 0)) >> 20}


def once():
    tracker.clear_device_cache()
    return tracker.track(models, tile_size=scene.tile_size)


def run(n, label):
    rows, tracks = [], None
    for _ in range(n):
        before = counters()
        t0 = time.perf_counter()
        tracks = once()
        wall = 1e3 * (time.perf_counter() - t0)
        after = counters()
        row = {"wall": round(wall, 1)}
        row.update({k: round(v, 1) for k, v in tracker.last_run.get("host_ms", {}).items()})
        row.update({k: after[k] - before[k] for k in ("dev_alloc", "dev_free", "host_alloc", "host_free")})
        row.update({"dev_reserved_mb": after["dev_reserved_mb"], "host_mb": after["host_mb"]})
        rows.append(row)
    everyone = [rows]
    if dist is not None:
        everyone = [None] * world
        dist.all_gather_object(everyone, rows)
    if rank == 0:
        for i in range(n):
            walls = [r[i]["wall"] for r in everyone]
            slow = max(range(world), key=lambda k: everyone[k][i]["session"] + everyone[k][i]["enqueue"])
            print(json.dumps({"phase": label, "call": i, "wall_max": max(walls), "wall_min": min(walls), "rank0": everyone[0][i],
                              "slowest_host_rank": slow, "slowest_host": everyone[slow][i]}), flush=True)
    return tracks


tracks = run(3, "warmup")
torch.cuda.synchronize()
if dist is not None:
    dist.barrier()
gc.collect()
gc.disable()
tracks = run(calls, "gc_disabled")
gc.enable()
tracks = run(max(3, calls // 2), "gc_enabled")
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
