#!/bin/bash
# Sweep the streaming-mode batch/slot plan on the bench workload (tuning aid): tools/sweep_slots.sh 1:1000 4:250 ...
# Prints G updates/s, ms per track, e2e G/s and the per-kernel milliseconds per track (CUDA events per launch).
for cfg in "$@"; do
  s=${cfg%%:*}; b=${cfg##*:}
  GB_STREAM_SLOTS=$s GB_STREAM_BATCH=$b timeout 300 python bench.py --no-cpu 2>&1 | tail -1 | CFG="slots=$s batch=$b" python -c '
import json, os, sys
d = json.loads(sys.stdin.read())
k = {n[2:6]: round(v["ms_per_track"], 1) for n, v in d["roofline"].get("kernels", {}).items()}
print(os.environ["CFG"], round(d["value"] / 1e9, 2), "G/s", round(d["ms_per_step"], 1), "ms  e2e", round(d["e2e"]["value"] / 1e9, 2), k)'
done
