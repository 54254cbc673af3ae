# parity tests of the wave paths (incl. full size) with the default library, then timings
python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "wave or fused" > gpurun_out/pytest_fusedq.log 2>&1; tail -2 gpurun_out/pytest_fusedq.log
B="python bench.py --steps 30 --warmup 3 --skip-e2e --skip-cpu --skip-count --skip-sw"
for rep in 1 2; do $B > gpurun_out/bench_var.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_var.log").read().strip().splitlines()[-1]); print("default", round(d["value"],1), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4), d["parity_spot_check"]["bit_exact"], d["roofline"].get("replay_fraction"))
PY
done
