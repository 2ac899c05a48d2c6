"""BASELINE.json configs 3 and 4 at their full sizes through the public API (one GPU).

    python tools/full_size_configs.py [3] [4] [5]

config 3: CylindricalMotion with uncertain elevation (dem_sigma = 1 m), 2 observers, 10 000 points x 10 000 particles x
          100 frames of 4288 x 2848 (1e10 particle-updates; the second station sees the same ground rolled by 180 deg
          and starts one frame later, so templates are staggered).
config 4: large-tile stress, 31 x 31 template and ~100 px search windows, 1 000 points x 100 000 particles x 50 frames
          (5e9 particle-updates; every point's particles span 140 CTAs of k_s3 / k_s4p).
config 5: the whole-box workload on ONE GPU: 100 000 points x 10 000 particles x 365 frames (3.65e11 particle-updates).  Its work
          buffers (~300 GB) exceed one B200, so the Tracker advances consecutive blocks of points (see Tracker._track_local);
          on 8 GPUs every rank holds its 12 500 points at once.  Not in the default list (about half a minute of GPU time).

The oracle cannot run these sizes; what is checked are size-independent properties (the same ones
tests/test_gpu_full_size.py checks for config 2): no failed point, finite moments, every point recovers the synthetic
velocity, uncertainties shrink.  Prints one JSON line per config with the end-to-end rate of the call.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from glimpse_b200 import synthetic  # noqa: E402

IMGSZ = (4288, 2848)


def second_station(scene):
    """Same place, rolled 180 deg (frames flipped both ways), radial-only distortion, first image one frame later."""
    first = scene.observers[0]
    frames, cams, dts = [], [], []
    for t in range(1, len(first.frames)):
        frames.append(np.ascontiguousarray(first.frames[t][::-1, ::-1]))
        vec = first.cams[t].copy()
        vec[3:6] = (0.0, -90.0, 180.0)
        vec[18:20] = 0.0
        cams.append(vec)
        dts.append(first.datetimes[t])
    scene.observers.append(synthetic.ObserverScene(frames, np.array(cams), dts, sigma=0.4))
    return scene


def scene_for(config):
    if config == 3:
        scene = synthetic.nadir_scene(seed=3, n_points=10000, n_particles=10000, n_frames=100, imgsz=IMGSZ, kind="cylindrical",
                                      velocity_sigma=0.2, margin_px=200)
        return second_station(scene)
    if config == 4:
        return synthetic.nadir_scene(seed=4, n_points=1000, n_particles=100000, n_frames=50, imgsz=IMGSZ, tile_size=(31, 31),
                                     velocity_sigma=0.3, margin_px=300)
    if config == 5:
        return synthetic.nadir_scene(seed=5, n_points=100000, n_particles=10000, n_frames=365, imgsz=IMGSZ, velocity_sigma=0.2,
                                     shift_px=(1, 0), margin_px=450)
    raise SystemExit("config must be 3, 4 or 5")


def check_and_report(config, scene, tracks, seconds, stats):
    P, T = len(scene.points), len(scene.datetimes)
    N = scene.n_particles
    errors = [e for e in tracks.errors if e is not None]
    kinds = sorted({type(e).__name__ + ": " + str(e).split(" (track")[0] for e in errors})
    ok = np.array([e is None for e in tracks.errors])
    v = tracks.vxyz[:, -1]
    dv = np.abs(v[ok, 0] - scene.truth_velocity[0])
    travelled = tracks.means[ok, -1, 0] - tracks.means[ok, 0, 0]
    out = {
        "config": config, "points": P, "particles": N, "frames": T, "observers": len(scene.observers),
        "tile_size": list(scene.tile_size), "seconds_e2e": seconds, "particle_updates_per_s_e2e": P * N * T / seconds,
        "failed_points": int((~ok).sum()), "failure_kinds": kinds,
        "finite": bool(np.isfinite(tracks.means[ok]).all() and np.isfinite(tracks.sigmas[ok]).all()),
        "median_abs_velocity_error": float(np.median(dv)), "max_abs_velocity_error": float(dv.max()),
        "max_rel_travel_error": float(np.abs(travelled / (scene.truth_velocity[0] * (T - 1)) - 1).max()),
        "median_sigma_vx_first_update": float(np.median(tracks.sigmas[ok, 1, 3])),
        "median_sigma_vx_last": float(np.median(tracks.sigmas[ok, -1, 3])),
        "search_window_px": {k: float(np.percentile(stats["window_width"], q)) for k, q in
                             (("median_w", 50), ("p90_w", 90), ("p99_w", 99), ("max_w", 100))},
        "scratch_gb": stats["plan"]["scratch_bytes"] / 1e9, "kernel_launches": stats["kernel_launches"],
        "sessions": stats.get("sessions", 1), "points_per_session": stats["plan"]["stream_batch"] * stats["plan"]["stream_slots"],
    }
    print(json.dumps(out), flush=True)
    return out


def run(config):
    import torch

    import glimpse_b200 as gb

    rank, world = 0, 1
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:  # torchrun: points sharded over the GPUs of the box
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
        rank, world = dist.get_rank(), dist.get_world_size()
    t0 = time.perf_counter()
    scene = scene_for(config)
    observers, models = synthetic.build(scene, gb)
    print(f"config {config}: scene built in {time.perf_counter() - t0:.1f} s", file=sys.stderr, flush=True)
    tracker = gb.Tracker(observers, seed=20260100 + config)
    best = None
    for rep in range(1 if (config == 5 and world == 1) else 2):  # the second call finds the frames on the device and the allocator warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tracks = tracker.track(models, tile_size=scene.tile_size)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        print(f"config {config}: track call {rep} took {dt:.2f} s", file=sys.stderr, flush=True)
    if rank != 0:
        return None
    out = check_and_report(config, scene, tracks, best, tracker.last_run)
    if world > 1:
        print(json.dumps({"config": config, "n_gpus": world, "host_ms_rank0": tracker.last_run.get("host_ms")}), flush=True)
    return out


if __name__ == "__main__":
    for c in [int(a) for a in sys.argv[1:]] or [3, 4]:
        run(c)
