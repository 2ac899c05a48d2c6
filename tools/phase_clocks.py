"""Per-phase clock64() breakdown of k_s2_surface (every CTA of every update), for tuning.

    python tools/phase_clocks.py [--points 1000] [--frames 60]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=592)
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--particles", type=int, default=10000)
    ap.add_argument("--mode", default="stream", choices=["stream"])
    args = ap.parse_args()
    import torch

    import glimpse_b200 as gb
    from glimpse_b200 import _lib, synthetic
    from glimpse_b200.session import Session

    bench.WORKLOAD["n_particles"] = args.particles
    scene = bench.build_scene(args.points, args.frames)
    observers, models = synthetic.build(scene, gb)
    tracker = gb.Tracker(observers, seed=1)
    dts = tracker.datetimes
    index = np.array([[-1 if v is None else int(v) for v in row] for row in tracker.match_datetimes(dts)], dtype=np.int32)
    taus = np.ones(len(dts) - 1)
    s = Session(tracker, models, index, taus, scene.tile_size, np.ones((args.points, 1), dtype=bool))
    s.init(0)
    clocks = torch.zeros((args.points, 16), dtype=torch.int64, device=s.device)
    io = _lib.gb_stage_io()
    io.dump_clocks = clocks.data_ptr()
    for t in range(1, args.frames):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s.step(t, io)
        b.record()
        torch.cuda.synchronize()
        c = clocks.cpu().numpy().astype(np.float64)
        if args.mode == "stream":
            # k_s2_surface: 0 entry, 1 window known, 2 after LUT, 3 after median, 4 after SSD, 5 after Hermite, 6 exit; 7/8 window, 9 SM
            if t in (1, 2, args.frames // 2, args.frames - 1):
                ok = c[:, 6] > 0
                size = c[ok, 7]
                tot = c[ok, 6] - c[ok, 0]
                print(f"t={t}: update {a.elapsed_time(b) * 1e3:.0f} us; k_s2 CTAs {ok.sum()}, window median {np.median(size):.0f} max {size.max():.0f}; "
                      f"CTA clocks median {np.median(tot):.0f} p90 {np.percentile(tot, 90):.0f} max {tot.max():.0f}")
                names = ["box", "load+hist+LUT", "median", "SSD", "Hermite", "store"]
                for lo, hi in ((0, 40), (40, 64), (64, 100), (100, 400)):
                    m = ok.copy()
                    m[ok] = (size >= lo) & (size < hi)
                    if not m.any():
                        continue
                    parts = [np.mean(c[m, k + 1] - c[m, k]) for k in range(6)]
                    print(f"    windows [{lo},{hi}) n={m.sum():4d} total {np.mean(c[m, 6] - c[m, 0]):8.0f} clk: " +
                          ", ".join(f"{n} {v:.0f}" for n, v in zip(names, parts)))
            continue


if __name__ == "__main__":
    main()
