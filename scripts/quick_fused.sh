python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "wave or fused" > gpurun_out/pytest_fusedq.log 2>&1; tail -3 gpurun_out/pytest_fusedq.log
python bench.py --steps 20 --warmup 3 --skip-e2e --skip-cpu --skip-count > gpurun_out/bench_fusedq.log 2>&1; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_fusedq.log").read().strip().splitlines()[-1]); print("default", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("replay_fraction"), d["parity_spot_check"])
PY
for f in gams_b200/csrc/build/variants/*.so; do [ -e "$f" ] || continue; GAMS_CUDA_LIB=$PWD/$f python bench.py --steps 20 --warmup 3 --skip-e2e --skip-cpu --skip-count > gpurun_out/bench_var.log 2>&1; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_var.log").read().strip().splitlines()[-1]); print("$f", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["parity_spot_check"])
PY
done
