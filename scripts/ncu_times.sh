#!/bin/bash
# usage (under gpurun): bash scripts/ncu_times.sh "kernel regex" COUNT cmd... -> launch times of the matching kernels
K=$1; C=$2; shift; shift
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c $C --csv --log-file /tmp/l.csv "$@" > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("/tmp/l.csv")) if len(r)>10]
h=rows[0]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[kn][:45],[]).append(float(r[mv].replace(",",""))/1e6)
for k,v in agg.items(): print(f"{k:45s} n={len(v)} sum={sum(v):.3f} ms  last={v[-1]:.3f}")
PY
