#!/bin/bash
# usage (under gpurun): bash scripts/r2_profile.sh TAG -> gpurun_out/{launches,ncu_full_summary,ncu_key_metrics}_TAG.*, the same with _gz
# (each ncu pass only after the same command has exited 0 without ncu; the .ncu-rep files stay on the box: gpurun_out is capped at 64 MiB)
TAG=$1
CMD="python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --skip-sw"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gc_wave_fused|gc_window_kernel|zscore_indep|count_overlaps_staged|seg_query_bounds|seg_items|cw_tile|part_sort|part_lookup|part_unpermute" -c 40 -o /tmp/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -1 gpurun_out/ncu_full_$TAG.log
python profiles/summarize_ncu.py /tmp/prof_$TAG.ncu-rep gpurun_out/ncu_full_summary_$TAG.csv
python profiles/ncu_key.py /tmp/prof_$TAG.ncu-rep > gpurun_out/ncu_key_metrics_$TAG.txt
python scripts/gz_probe.py > gpurun_out/plain_gz_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_gz_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"inflate_|gunzip_kernel|crc32_" -c 16 --csv --log-file gpurun_out/launches_gz_$TAG.csv python scripts/gz_probe.py > gpurun_out/ncu_gz_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"inflate_split|inflate_tokens|inflate_resolve|crc32_pieces" -c 4 -o gpurun_out/prof_gz_$TAG python scripts/gz_probe.py > gpurun_out/ncu_gzfull_$TAG.log 2>&1
tail -1 gpurun_out/ncu_gzfull_$TAG.log
python profiles/summarize_ncu.py gpurun_out/prof_gz_$TAG.ncu-rep gpurun_out/ncu_full_summary_gz_$TAG.csv
python profiles/ncu_key.py gpurun_out/prof_gz_$TAG.ncu-rep > gpurun_out/ncu_key_metrics_gz_$TAG.txt
