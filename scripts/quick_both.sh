# quick A/B on one box: parity tests of the count and wave paths, then the default library and every variant
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "count or wave or fused" > gpurun_out/pytest_q.log 2>&1; tail -3 gpurun_out/pytest_q.log
B="python bench.py --steps 30 --warmup 3 --skip-e2e --skip-cpu --skip-sw"
for rep in 1 2; do
for f in default gams_b200/csrc/build/variants/*.so; do [ "$f" = default ] || [ -e "$f" ] || continue
if [ "$f" = default ]; then $B > gpurun_out/bench_var.log 2>&1; else GAMS_CUDA_LIB=$PWD/$f $B > gpurun_out/bench_var.log 2>&1; fi
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_var.log").read().strip().splitlines()[-1]); o=d["overlap"]["grouped_by_ctg"]; a=d["overlap"]["arbitrary_order"]
print("$f", "wave", round(d["ms_per_step"],4), round(d["roofline"]["frac"],4), d["parity_spot_check"]["bit_exact"], "| count", round(o["ms_per_step"],4), round(o["roofline"]["frac"],4), o["bit_exact_sample"], "| any-order", round(a["ms_per_step"],3))
PY
done; done
