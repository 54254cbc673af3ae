#!/bin/bash
# usage (under gpurun): bash scripts/r2_prof.sh TAG REGEX [probe args]  -> ncu --set full of the kernels matching REGEX in count_probe.py
TAG=$1; RE=$2; shift; shift
timeout 200 python scripts/count_probe.py "$@" 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$RE" -c 4 -o gpurun_out/prof_$TAG python scripts/count_probe.py "$@" > gpurun_out/ncu_prof_$TAG.log 2>&1
tail -1 gpurun_out/ncu_prof_$TAG.log
