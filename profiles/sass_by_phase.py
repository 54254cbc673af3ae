#!/usr/bin/env python3
"""Per-phase dynamic instruction counts of a kernel: joins the SASS page of an ncu report with nvdisasm line info and
groups source lines by the `// ---- ` phase comments of the .cu file.
usage: sass_by_phase.py report.ncu-rep file.cubin kernel_substring source.cu n_tiles"""
import collections, csv, re, subprocess, sys

rep, cubin, kname, src, ntile = sys.argv[1:6]
ntile = float(ntile)
marks = [(1, "0 setup")]
for i, ln in enumerate(open(src), 1):
    m = re.match(r"\s*// ---- (.{0,28})", ln)
    if m and "kernel" not in ln:
        marks.append((i, f"{len(marks)} {m.group(1).strip()}"))
kstart = next(i for i, ln in enumerate(open(src), 1) if "__global__" in ln)
dis = subprocess.run(["nvdisasm", "-c", "--print-line-info", cubin], capture_output=True, text=True).stdout
lines, ops, inside, cur = [], [], False, None
for ln in dis.splitlines():
    if ln.startswith(".text."):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", ln)
    if m:
        lines.append(cur)
        ops.append(m.group(1))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ix = {n: i for i, n in enumerate(h)}
body = [r for r in rows[hi + 1:] if len(r) >= len(h)]
assert len(body) == len(lines), (len(body), len(lines))
base = src.split("/")[-1]


def phase(l, prev):
    if l is None:
        return prev
    f, n = l
    if f != base or n < kstart:
        return prev  # inlined helper: belongs to the phase that called it
    name = marks[0][1]
    for ln, nm in marks:
        if n >= ln:
            name = nm
    return name


agg, samp, byop = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
tot = tots = 0
prev = "0 setup"
for r, l, o in zip(body, lines, ops):
    i, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    prev = p = phase(l, prev)
    agg[p] += i
    samp[p] += s
    tot += i
    tots += s
    t = o.split()
    byop[p][t[1] if t[0].startswith("@") else t[0]] += i
print(f"total warp instructions {tot}  ({tot / ntile / 8:.0f} per thread per tile), samples {tots}")
for p in sorted(agg, key=lambda x: int(x.split()[0])):
    print(f"{p:32s} inst {100 * agg[p] / tot:5.1f}%  samples {100 * samp[p] / tots:5.1f}%  per-thread {agg[p] / ntile / 8:6.1f}   " +
          ", ".join(f"{k} {v / ntile / 8:.0f}" for k, v in byop[p].most_common(10)))
