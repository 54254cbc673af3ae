python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "wave or fused" > gpurun_out/pytest_fusedq.log 2>&1; tail -3 gpurun_out/pytest_fusedq.log
B="python bench.py --steps 30 --warmup 3 --skip-e2e --skip-cpu --skip-count"
for rep in 1 2; do
for f in default gams_b200/csrc/build/variants/*.so; do [ "$f" = default ] || [ -e "$f" ] || continue
if [ "$f" = default ]; then $B > gpurun_out/bench_var.log 2>&1; else GAMS_CUDA_LIB=$PWD/$f $B > gpurun_out/bench_var.log 2>&1; fi
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_var.log").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4), d["parity_spot_check"]["bit_exact"])
PY
done; done
