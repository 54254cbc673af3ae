# chained gz path in pieces: tests, then ms per call for PIECES[:STREAMS]
python -m pytest tests/test_round2_gpu.py -m gpu -x -q -k "gz or wave" > gpurun_out/pytest_r2h_gz.log 2>&1; echo "gz/wave tests: $(tail -1 gpurun_out/pytest_r2h_gz.log)"
python scripts/trace_gz_wave.py "$@" > gpurun_out/trace_gz_wave_r2h.log 2>&1; tail -12 gpurun_out/trace_gz_wave_r2h.log
