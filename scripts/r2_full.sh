#!/bin/bash
# usage (under gpurun): bash scripts/r2_full.sh TAG -> the whole -m gpu suite, smoke, the default bench line and the reference arm
TAG=${1:-x}
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err || { tail -20 gpurun_out/bench_$TAG.err; }
tail -c 600 gpurun_out/bench_$TAG.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; tail -c 900 gpurun_out/bench_ref_$TAG.log
