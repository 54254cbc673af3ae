# influence != 1: thread-per-segment certified walk — tests, then the wave step at --influence 0.5 with the new and the old kernels
timeout 900 python -m pytest tests/test_round2_gpu.py -x -q -m gpu -k "thread_walk" > gpurun_out/pytest_r2j_walk.log 2>&1; echo "thread-walk test: $(tail -1 gpurun_out/pytest_r2j_walk.log)"
timeout 1200 python -m pytest tests -x -q -m gpu -k "influence or thresholding or wave" > gpurun_out/pytest_r2j_infl.log 2>&1; echo "influence / thresholding / wave tests: $(tail -1 gpurun_out/pytest_r2j_infl.log)"
B="python bench.py --steps 5 --warmup 3 --influence 0.5 --skip-count --skip-sw --skip-cpu --skip-gz --skip-frontend --skip-e2e"
for v in new old; do
if [ $v = old ]; then GAMS_CUDA_ZWALK_WARPS=1 timeout 600 $B > gpurun_out/bench_walk_$v.log 2>gpurun_out/bench_walk_$v.err; else timeout 600 $B > gpurun_out/bench_walk_$v.log 2>gpurun_out/bench_walk_$v.err; fi
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_walk_$v.log").read().strip().splitlines()[-1]); print("$v", "ms_per_pass", round(d.get("ms_per_pass", d["ms_per_step"]),4), "value", round(d["value"],1), d["parity_spot_check"], d.get("path"))
PY
done
