#!/bin/bash
# usage (under gpurun): bash scripts/r2_count.sh TAG  -> tests of the overlap-count paths, a count-only bench line and its ncu launch list
TAG=${1:-x}
timeout 300 python -m pytest tests/test_round2_gpu.py -x -q --timeout 120 2>&1 | tail -15
timeout 400 python -m pytest tests/test_parity_gpu.py -x -q --timeout 120 -k "count or locate or thresholding or zscore or fused or parse" 2>&1 | tail -5
B="timeout 300 python bench.py --steps 5 --warmup 3 --skip-e2e --skip-cpu --skip-sw --skip-gz"
$B > gpurun_out/bench_count_$TAG.log 2> gpurun_out/bench_count_$TAG.err || { tail -20 gpurun_out/bench_count_$TAG.err; exit 1; }
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_count_$TAG.log").read().strip().splitlines()[-1]); o=d["overlap"]
for k in ("grouped_by_ctg","arbitrary_order","sliding_windows"):
    print(k, round(o[k]["ms_per_step"],4), "ms  frac", round(o[k]["roofline"]["frac"],3), o[k].get("bit_exact_sample"), "launches", o[k]["gpu_launches"])
print("wave", d["ms_per_step"], d["roofline"]["frac"])
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"part_|count_overlaps|seg_" -c 200 --csv --log-file gpurun_out/launches_count_$TAG.csv $B > gpurun_out/ncu_count_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"part_sort|part_lookup|part_unpermute" -c 3 -o gpurun_out/prof_count_$TAG $B > gpurun_out/ncu_full_count_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_count_$TAG.csv")) if len(r)>10]
h=rows[0]; kn=h.index("Kernel Name"); mn=h.index("Metric Name"); mv=h.index("Metric Value"); idc=h.index("ID")
agg={}
for r in rows[1:]:
    agg.setdefault((r[idc],r[kn]),{})[r[mn]]=float(r[mv].replace(",",""))
seen={}
for (i,k),m in agg.items():
    seen.setdefault(k,[]).append(m)
for k,ms in seen.items():
    m=ms[-1]
    print(f"{k[:60]:60s} n={len(ms):3d} last: {m.get('gpu__time_duration.sum',0)/1e3:9.1f} us  rd {m.get('dram__bytes_read.sum',0)/1e9:7.3f} GB  wr {m.get('dram__bytes_write.sum',0)/1e9:7.3f} GB")
PY
