python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "count_windows" > gpurun_out/pytest_q.log 2>&1; tail -1 gpurun_out/pytest_q.log
B="python bench.py --steps 20 --warmup 3 --skip-e2e --skip-cpu --skip-sw"
for f in default gams_b200/csrc/build/variants/*.so; do [ "$f" = default ] || [ -e "$f" ] || continue
if [ "$f" = default ]; then $B > gpurun_out/bench_var.log 2>&1; else GAMS_CUDA_LIB=$PWD/$f $B > gpurun_out/bench_var.log 2>&1; fi
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_var.log").read().strip().splitlines()[-1]); o=d["overlap"]["grouped_by_ctg"]; a=d["overlap"]["sliding_windows"]
print("$f", "| count", round(o["ms_per_step"],4), o["bit_exact_sample"], "| windows", round(a["ms_per_step"],4), round(a["roofline"]["frac"],4), a["bit_exact_sample"])
PY
done
