python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "wave or fused" > gpurun_out/pytest_fused2.log 2>&1; tail -5 gpurun_out/pytest_fused2.log
B="python bench.py --steps 20 --warmup 3 --skip-e2e --skip-cpu --skip-count"
$B > gpurun_out/bench_fused2.log 2>&1; python - <<'PY'
import json
for f in ["gpurun_out/bench_fused2.log"]:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["parity_spot_check"])
PY
for m in 5 6; do GAMS_CUDA_LIB=$PWD/gams_b200/csrc/build/variants/lib_ctas$m.so $B > gpurun_out/bench_fused2_c$m.log 2>&1; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_fused2_c$m.log").read().strip().splitlines()[-1]); print("ctas$m", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["parity_spot_check"])
PY
done
