#!/bin/bash
# usage (under gpurun): bash scripts/r2_zseg.sh TAG -> influence != 1 tests with the default segment length and with boundaries every 300 / 137 windows; wave step at influence 0.5
TAG=${1:-x}
K="influence or thresholding or zscore or walk or wave_synthetic or wave_gz"
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 -k "$K" 2>&1 | tail -3
GAMS_CUDA_ZSEG=300 timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 -k "$K" 2>&1 | tail -3
GAMS_CUDA_ZSEG=137 timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu --timeout 600 -k "influence or thresholding" 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --influence 0.5 --skip-count --skip-sw --skip-cpu --skip-gz --skip-frontend > gpurun_out/bench_infl_$TAG.log 2> gpurun_out/bench_infl_$TAG.err || tail -5 gpurun_out/bench_infl_$TAG.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_infl_$TAG.log").read().strip().splitlines() if l.startswith("{")][-1])
print("wave step influence 0.5:", d["ms_per_pass"], "ms per pass,", d["value"], "Gbp/s; e2e", d["e2e"] and d["e2e"]["ms_per_step"], "path", d.get("path"), d.get("parity_spot_check"))
PY
timeout 300 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"zscore_seg|zscore_seq" -c 4 --csv --log-file gpurun_out/launches_zseg_$TAG.csv python bench.py --steps 1 --warmup 1 --passes 1 --influence 0.5 --skip-count --skip-sw --skip-cpu --skip-gz --skip-frontend --skip-e2e > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_zseg_$TAG.csv | cut -d, -f5,13,15 | cut -c1-160
