#!/bin/bash
# usage (under gpurun): bash scripts/r2_profile_gz.sh TAG -> launch list and ncu --set full summaries of the gunzip kernels (scripts/gz_probe.py)
TAG=$1
python scripts/gz_probe.py > gpurun_out/plain_gz_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_gz_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"inflate_|gunzip_kernel|crc32_" -c 16 --csv --log-file gpurun_out/launches_gz_$TAG.csv python scripts/gz_probe.py > gpurun_out/ncu_gz_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"inflate_split|inflate_tokens|inflate_resolve|crc32_pieces" -c 4 -o gpurun_out/prof_gz_$TAG python scripts/gz_probe.py > gpurun_out/ncu_gzfull_$TAG.log 2>&1
tail -1 gpurun_out/ncu_gzfull_$TAG.log
python profiles/summarize_ncu.py gpurun_out/prof_gz_$TAG.ncu-rep gpurun_out/ncu_full_summary_gz_$TAG.csv
python profiles/ncu_key.py gpurun_out/prof_gz_$TAG.ncu-rep > gpurun_out/ncu_key_metrics_gz_$TAG.txt
cat gpurun_out/plain_gz_$TAG.log | tail -3
