#!/bin/bash
# usage (under gpurun): bash scripts/r2_zseg_sweep.sh -> wave step at influence 0.5 for several segment lengths of the walk
for Z in 16384 8192 4096 2048 1024; do
GAMS_CUDA_ZSEG=$Z timeout 600 python bench.py --steps 5 --warmup 3 --influence 0.5 --skip-count --skip-sw --skip-cpu --skip-gz --skip-frontend --skip-e2e > gpurun_out/bench_zseg_$Z.log 2>/dev/null
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_zseg_$Z.log").read().strip().splitlines() if l.startswith("{")][-1])
print("seg $Z: wave step influence 0.5:", round(d["ms_per_pass"],3), "ms per pass", d.get("parity_spot_check",{}).get("bit_exact"))
PY
done
