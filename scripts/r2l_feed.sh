# gunzip_batch in pieces: tests, then the bench line's feed record (device-resident decoders + the host entry point)
timeout 600 python -m pytest tests -x -q -m gpu -k "gunzip or gz" > gpurun_out/pytest_r2l_gz.log 2>&1; echo "gunzip / gz tests: $(tail -1 gpurun_out/pytest_r2l_gz.log)"
timeout 600 python bench.py --steps 3 --warmup 3 --skip-count --skip-sw --skip-frontend --skip-cpu > gpurun_out/bench_r2l.log 2> gpurun_out/bench_r2l.err || tail -20 gpurun_out/bench_r2l.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_r2l.log").read().strip().splitlines() if l.startswith("{")][-1])
f=d["feed"]; print("feed device", round(f["ms_per_step"],2), "ms; host e2e", f["e2e"])
print("e2e_gz", round(d["e2e_gz"]["ms_per_step"],2), "peaks", round(d["e2e_gz"]["peaks_only"]["ms_per_step"],2))
PY
for P in 1 3; do GAMS_CUDA_GZ_PIECES=$P timeout 600 python bench.py --steps 3 --warmup 3 --skip-count --skip-sw --skip-frontend --skip-cpu > gpurun_out/bench_r2l_p$P.log 2>/dev/null; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r2l_p$P.log").read().strip().splitlines() if l.startswith("{")][-1])
print("pieces $P: host gunzip e2e", round(d["feed"]["e2e"]["ms_per_step"],2), "ms", d["feed"]["e2e"]["bit_exact"], "; gz -> windows", round(d["e2e_gz"]["ms_per_step"],2))
PY
done
