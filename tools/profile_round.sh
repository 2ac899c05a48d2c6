#!/bin/bash
# One gpurun call that collects everything a round's profiles/ needs (about 4 minutes of box time on one GPU):
#
#   gpurun --timeout 600 -- 'bash tools/profile_round.sh r2'
#
#   gpurun_out/<tag>_bench.json          the bench line (value, e2e, roofline, cpu_baseline, configs, strong, result_crc), not under a profiler
#   gpurun_out/<tag>_launches.csv        ncu launch list of a short bench run (durations, instructions, DRAM bytes, pipe utilisation per launch)
#   gpurun_out/<tag>_<kernel>.ncu-rep    ncu --set full of one launch of each hot kernel (read with `ncu -i ... --page raw --csv`)
#   gpurun_out/<tag>_<kernel>.txt        its details page as text
# Numbers printed under ncu are never bench values: clocks are not locked (--clock-control none), caches are flushed per launch.
set -u
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
metrics=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
timeout 240 ncu --metrics $metrics --clock-control none -c 300 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --frames 40 --steps 1 --warmup 1 --no-cpu --no-side > $out/${tag}_ncu_bench.log 2>&1
for k in k_s4p_resample_propagate k_s2_surface k_s3_weights; do
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 1 -f -o $out/${tag}_$k \
    python bench.py --frames 30 --steps 1 --warmup 1 --no-cpu --no-side > $out/${tag}_ncu_$k.log 2>&1
  ncu -i $out/${tag}_$k.ncu-rep --page details > $out/${tag}_$k.txt 2>/dev/null
done
tail -c 300 $out/${tag}_bench.json
