"""`gams locate --count -f file` text in, text out, at scale: 2e7 query lines against 1e8 SNP intervals on the bench genome."""
import sys, time
import numpy as np
from gams_b200 import api, synth
import bench
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
g_off, ctgs = bench.layout(int(3.1e9))
lens = np.diff(g_off.astype(np.int64)); n_ctg = lens.size
chr_of = np.array([c[0] for c in ctgs]); names = [f"chr{k + 1}" for k in range(int(chr_of.max()) + 1)]
cs = np.array([c[1] for c in ctgs], dtype=np.uint32); ce = np.array([c[2] + 1 for c in ctgs], dtype=np.uint32)   # stop exclusive
pos = {k: int(ce[chr_of == k].max()) for k in range(len(names))}
chr_off = np.concatenate(([0], np.cumsum(np.bincount(chr_of, minlength=len(names))))).astype(np.uint64)
rng = np.random.default_rng(3)
per = 100_000_000 // n_ctg
st = np.concatenate([np.sort(rng.integers(int(cs[k]), int(ce[k]), per)).astype(np.uint32) for k in range(n_ctg)])
sp = st + 1
iv_off = (np.arange(n_ctg + 1, dtype=np.uint64) * np.uint64(per))
m = 1_000_000
c = rng.integers(0, len(names), m); chr_len = np.array([pos[k] - 1 for k in range(len(names))])
s = (rng.random(m) * chr_len[c]).astype(np.int64) + 1; e = np.minimum(s + rng.integers(0, 10_000, m), chr_len[c])
text = ("\n".join(f"{names[a]}:{b}-{d}" for a, b, d in zip(c.tolist(), s.tolist(), e.tolist())) + "\n").encode() * (n // m)
print(f"{len(text) / 1e6:.0f} MB of query text, {n} lines; {st.size} intervals on {n_ctg} ctgs", flush=True)
ctx = api.Context([0])
api.locate_count_text(ctx, text[:100000], names, cs, ce, chr_off, st, sp, iv_off)
for _ in range(2):
    t = time.perf_counter(); out, nl, npr, nb = api.locate_count_text(ctx, text, names, cs, ce, chr_off, st, sp, iv_off); dt = time.perf_counter() - t
print(f"device: {dt * 1e3:.0f} ms end to end (text + both indexes up, {len(out) / 1e6:.0f} MB of text down) = {nl / dt / 1e6:.1f} M lines/s; printed {npr}, handed back {nb}")
from oracle import oracle as orc
k = 2000
lines = text.decode().splitlines()[:k]
got = out.decode().splitlines()[:k]
seg_of = {}
ok = True
for ln, g in zip(lines, got):
    rg = orc.range_from_str(ln); ci = names.index(rg[0])
    j = int(np.searchsorted(ce[int(chr_off[ci]):int(chr_off[ci + 1])], rg[2], side="right")) + int(chr_off[ci])
    ref = orc.lapper_count(st[int(iv_off[j]):int(iv_off[j + 1])], sp[int(iv_off[j]):int(iv_off[j + 1])], rg[2], rg[3])
    ok &= g == f"{ln}\t{ref}"
print("first", k, "output lines identical to count_rg:", ok)
