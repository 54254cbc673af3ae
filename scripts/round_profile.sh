#!/bin/bash
# usage (under gpurun): bash scripts/round_profile.sh TAG   -> gpurun_out/{bench,launches,prof}_TAG.*
set -x
TAG=$1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err || exit 1
CMD="python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gc_wave_fused|gc_window_kernel|zscore_indep|count_overlaps_staged|count_overlaps_ranked|seg_query_bounds|seg_items|cw_tile|cw_desc" -c 36 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
