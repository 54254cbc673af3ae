#!/bin/bash
# usage (under gpurun): bash scripts/r2_gzprof.sh TAG [bases]  -> launch times and ncu --set full of the gunzip kernels in gz_probe.py
TAG=$1; N=${2:-3.1e9}
GAMS_GZ_PROBE_ONLY=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"inflate_|gunzip_kernel|crc32_" -c 12 --csv --log-file gpurun_out/launches_gz_$TAG.csv python scripts/gz_probe.py $N > gpurun_out/ncu_gz_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_gz_$TAG.csv")) if len(r)>10]
h=rows[0]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
for r in rows[1:]: print(f"{r[kn][:50]:50s} {float(r[mv].replace(',',''))/1e6:9.3f} ms")
PY
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"inflate_tokens|inflate_resolve" -c 2 -o gpurun_out/prof_gz_$TAG python scripts/gz_probe.py $N > gpurun_out/ncu_gzfull_$TAG.log 2>&1; tail -1 gpurun_out/ncu_gzfull_$TAG.log
