#!/bin/bash
# usage (under gpurun): bash scripts/r2_gzexp.sh -> launch times of the gunzip kernels for every build/variants/lib_*.so (timing experiments, results not checked)
for f in gams_b200/csrc/build/variants/lib_*.so; do
echo "== $f"
GAMS_GZ_PROBE_NOCHECK=1 GAMS_CUDA_LIB=$PWD/$f timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"inflate_tokens|inflate_resolve|inflate_split" -c 3 --csv --log-file /tmp/l.csv python scripts/gz_probe.py > /tmp/o.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("/tmp/l.csv")) if len(r)>10]
h=rows[0]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
for r in rows[1:]: print(f"  {r[kn][:40]:40s} {float(r[mv].replace(',',''))/1e6:9.3f} ms")
PY
done
