timeout 1200 python -m pytest tests -x -q -m gpu -k "influence or thresholding or wave" > gpurun_out/pytest_r2j_walk.log 2>&1; echo "influence tests: $(tail -1 gpurun_out/pytest_r2j_walk.log)"
B="python bench.py --steps 2 --warmup 1 --passes 1 --influence 0.5 --skip-count --skip-sw --skip-cpu --skip-gz --skip-frontend --skip-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_walk_r2j.csv $B > gpurun_out/ncu_walk.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_walk_r2j.csv")) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
agg=collections.defaultdict(list)
for r in rows[1:]:
    if "gams" in r[ki]:
        v=float(r[vi].replace(",","")); u=r[ui]
        v = v/1e6 if u in ("ns","nsecond") else (v/1e3 if u in ("us","usecond") else v)
        agg[r[ki][:60]].append(v)
for k,v in agg.items(): print(k, len(v), "launches, last", round(v[-1],3), "ms")
PY
python bench.py --steps 5 --warmup 3 --influence 0.5 --skip-count --skip-sw --skip-cpu --skip-gz --skip-frontend --skip-e2e > gpurun_out/bench_walk_new.log 2>/dev/null
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_walk_new.log").read().strip().splitlines() if l.startswith("{")][-1])
print("wave step influence 0.5:", round(d["ms_per_pass"],3), "ms per pass", d.get("parity_spot_check",{}))
PY
