#!/bin/bash
# usage (under gpurun): bash scripts/r2_gzprof.sh TAG [kernel regex] -> ncu --set full of the gunzip kernels in gz_probe.py (first launch of each)
TAG=$1; K=${2:-inflate_split|inflate_tokens|inflate_resolve}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -c 3 -o gpurun_out/prof_gz_$TAG python scripts/gz_probe.py > gpurun_out/ncu_gzfull_$TAG.log 2>&1; tail -2 gpurun_out/ncu_gzfull_$TAG.log
