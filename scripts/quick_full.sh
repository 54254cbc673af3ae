# whole GPU suite, then one default bench run
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; tail -3 gpurun_out/pytest_full.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; tail -n 1 gpurun_out/bench_full.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('wave', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e'])
for k,v in d['overlap'].items(): print(k, {a:b for a,b in v.items() if a in ('value','ms_per_step','bit_exact_sample','matches_device_path')}, v.get('roofline',{}).get('frac'))
print('sw', d['sw']); print('cpu', d['cpu_baseline'])
"
