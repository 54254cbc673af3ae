#!/bin/bash
# usage (under gpurun): bash scripts/r2k_profile.sh TAG -> the bench part of scripts/r2_profile.sh (launch list, ncu --set full summary and key
# metrics of the kernels the bench line reports a roofline for); the gunzip kernels did not change since r2f
TAG=$1
CMD="python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --skip-sw"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gc_wave_fused|gc_window_kernel|zscore_indep|count_overlaps_staged|seg_query_bounds|seg_items|cw_tile|part_sort|part_lookup|part_unpermute" -c 40 -o /tmp/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -1 gpurun_out/ncu_full_$TAG.log
python profiles/summarize_ncu.py /tmp/prof_$TAG.ncu-rep gpurun_out/ncu_full_summary_$TAG.csv
python profiles/ncu_key.py /tmp/prof_$TAG.ncu-rep > gpurun_out/ncu_key_metrics_$TAG.txt
