# round 2, session 3: chained gz path in pieces (tests, times per number of pieces) + A/B of the fused kernel at 512 threads per tile
python -m pytest tests/test_round2_gpu.py -m gpu -x -q -k "gz or wave" > gpurun_out/pytest_r2h_gz.log 2>&1; echo "gz/wave tests: $(tail -1 gpurun_out/pytest_r2h_gz.log)"
python scripts/trace_gz_wave.py 1 2:1 2 3 4 6 8 3:2 4:2 > gpurun_out/trace_gz_wave_r2h.log 2>&1; cat gpurun_out/trace_gz_wave_r2h.log | tail -8
for f in gams_b200/csrc/build/variants/*.so; do [ -e "$f" ] || continue
GAMS_CUDA_LIB=$PWD/$f python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "wave or fused" > gpurun_out/pytest_fusedq_$(basename $f).log 2>&1; echo "$f: $(tail -1 gpurun_out/pytest_fusedq_$(basename $f).log)"
done
B="python bench.py --steps 20 --warmup 3 --skip-e2e --skip-cpu --skip-count --skip-sw --skip-gz --skip-frontend"
for rep in 1 2; do
for f in default gams_b200/csrc/build/variants/*.so; do [ "$f" = default ] || [ -e "$f" ] || continue
if [ "$f" = default ]; then $B > gpurun_out/bench_var.log 2>&1; else GAMS_CUDA_LIB=$PWD/$f $B > gpurun_out/bench_var.log 2>&1; fi
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_var.log").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["ms_per_pass"],4), round(d["roofline"]["frac"],4), d["parity_spot_check"]["bit_exact"], d["roofline"].get("replay_fraction"), d["clocks"])
PY
done; done
