#!/bin/bash
# usage (under gpurun): bash scripts/r2_gz.sh TAG -> gunzip tests, gz_probe timings, launch list of the gunzip kernels
TAG=${1:-x}
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 -k "gunzip or gz" 2>&1 | tail -5
timeout 300 python scripts/gz_probe.py 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"inflate_|gunzip_kernel|crc32_" -c 14 --csv --log-file gpurun_out/launches_gz_$TAG.csv python scripts/gz_probe.py > gpurun_out/ncu_gz_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_gz_$TAG.csv")) if len(r)>10]
h=rows[0]; kn=h.index("Kernel Name"); mv=h.index("Metric Value"); g=h.index("Grid Size")
for r in rows[1:]: print(f"{r[kn][:40]:40s} {r[g]:14s} {float(r[mv].replace(',',''))/1e6:9.3f} ms")
PY
