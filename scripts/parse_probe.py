"""Throughput of gams_cuda_parse_ranges (host text in, arrays out) against Python's re on a sample: scripts/parse_probe.py [lines]"""
import sys, time
import numpy as np
from gams_b200 import api
from oracle import oracle as orc
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
rng = np.random.default_rng(1)
names = [f"chr{k}" for k in range(1, 25)]
c = rng.integers(1, 25, n); s = rng.integers(1, 250_000_000, n); e = s + rng.integers(0, 10_000, n)
t = time.perf_counter()
text = "\n".join(f"chr{a}:{b}-{d}" for a, b, d in zip(c[:1_000_000].tolist(), s[:1_000_000].tolist(), e[:1_000_000].tolist())) + "\n"
text = (text * (n // 1_000_000)).encode()
print(f"{len(text) / 1e6:.0f} MB of text, {n} lines (built in {time.perf_counter() - t:.1f} s)", flush=True)
ctx = api.Context([0])
api.parse_ranges(ctx, text[:1000], names)
for _ in range(2):
    t = time.perf_counter(); got = api.parse_ranges(ctx, text, names); dt = time.perf_counter() - t
print(f"device: {dt * 1e3:.1f} ms end to end (one call) = {n / dt / 1e6:.1f} M lines/s; handed back: {int((got['chr'] == -3).sum())}")
k = 200_000
lines = text[:int(got["line_off"][k])].decode().splitlines()
t = time.perf_counter(); ref = [orc.range_from_str(l) for l in lines]; dtc = time.perf_counter() - t
ok = all(r is not None and (int(got["start"][i]), int(got["end"][i])) == (r[2], r[3]) and names[int(got["chr"][i])] == r[0] for i, r in enumerate(ref))
print(f"python re on {k} lines: {k / dtc / 1e6:.2f} M lines/s; sample identical: {ok}")
