"""Per-phase clock64() breakdown of one k_step launch (rank-0 CTA of every point), for tuning.

    python tools/phase_clocks.py [--points 592] [--frames 6] [--cluster 0]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

NAMES = ["A evolve+test", "project+box allgather", "box,load,hist,LUT", "median highpass", "SSD", "hermite solve",
         "spline eval", "weights+scan", "W allgather", "child ranges E", "moments+allgather", "children write"]
# clock slots: 0 start, 1 after A, 2 after box allgather, 3 after LUT, 4 after median, 5 after SSD, 6 after hermite,
#              7 after eval, 8 after scan, 9 after W allgather, 10 after E, 11 after children, 12 end
ORDER = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8), (8, 9), (9, 10), (10, 11), (11, 12)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=592)
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--particles", type=int, default=10000)
    ap.add_argument("--mode", default="fused", choices=["fused", "stream"])
    args = ap.parse_args()
    import torch

    import glimpse_b200 as gb
    from glimpse_b200 import _lib, synthetic
    from glimpse_b200.session import Session

    bench.WORKLOAD["n_particles"] = args.particles
    scene = bench.build_scene(args.points, args.frames)
    observers, models = synthetic.build(scene, gb)
    tracker = gb.Tracker(observers, seed=1, cluster=args.cluster, mode=args.mode)
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
        ok = c[:, 12] > 0
        total = (c[ok, 12] - c[ok, 0]).mean()
        print(f"t={t}: launch {a.elapsed_time(b) * 1e3:.0f} us; rank-0 CTA mean {total:.0f} clk; plan cluster {s.plan.cluster}")
        for name, (i, j) in zip(NAMES, ORDER):
            d = (c[ok, j] - c[ok, i]).mean()
            print(f"    {name:26s} {d:9.0f} clk  {100 * d / total:5.1f} %")


if __name__ == "__main__":
    main()
