#!/bin/bash
# usage (under gpurun --gpus N): bash scripts/r2_multi.sh N TAG -> multi-device tests, then the bench line under torchrun
N=$1; TAG=${2:-x}
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q --timeout 300 -k "multi_device" 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 $BENCH_EXTRA > gpurun_out/bench_n${N}_$TAG.log 2> gpurun_out/bench_n${N}_$TAG.err || tail -20 gpurun_out/bench_n${N}_$TAG.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n${N}_$TAG.log").read().strip().splitlines() if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "e2e_gz", d["e2e_gz"] and (d["e2e_gz"]["value"], d["e2e_gz"]["ms_per_step"], d["e2e_gz"]["matches_device_path"]))
print(json.dumps(d["multi_gpu"], indent=1)[:3000])
PY
