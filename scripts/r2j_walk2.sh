# thread-per-segment walk: the new test, per-kernel times (ncu launch list of the influence-0.5 wave step), segment-length sweep
timeout 900 python -m pytest tests/test_round2_gpu.py -x -q -m gpu -k "thread_walk" > gpurun_out/pytest_r2j_walk.log 2>&1; echo "thread-walk test: $(tail -1 gpurun_out/pytest_r2j_walk.log)"
B="python bench.py --steps 2 --warmup 1 --passes 1 --influence 0.5 --skip-count --skip-sw --skip-cpu --skip-gz --skip-frontend --skip-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_walk_r2j.csv $B > gpurun_out/ncu_walk.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_walk_r2j.csv")) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
agg=collections.defaultdict(list)
for r in rows[1:]:
    if "gams" in r[ki] or "zscore" in r[ki] or "gc_wave" in r[ki]:
        v=float(r[vi].replace(",","")); u=r[ui]
        v = v/1e6 if u in ("ns","nsecond") else (v/1e3 if u in ("us","usecond") else v)
        agg[r[ki][:60]].append(v)
for k,v in agg.items(): print(k, len(v), "launches, last", round(v[-1],3), "ms")
PY
bash scripts/r2_zseg_sweep.sh
