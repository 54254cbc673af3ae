#!/usr/bin/env python3
"""Join an ncu SASS-level source page (ncu -i rep --page source --csv) with nvdisasm line info of the cubin, and
print per-source-line shares of executed instructions and stall samples.
usage: sass_by_line.py report.ncu-rep file.cubin kernel_substring [min_pct]"""
import csv, re, subprocess, sys, collections

rep, cubin, kname = sys.argv[1:4]
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
dis = subprocess.run(["nvdisasm", "-c", "--print-line-info", cubin], capture_output=True, text=True).stdout
# walk the function's text: remember the current line marker, assign it to each instruction in order
lines = []
inside = False
cur = None
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
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
        lines.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kname.split("IL")[0]], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ix = {n: i for i, n in enumerate(h)}
body = []
for r in rows[hi + 1:]:          # the first launch of the kernel only (a report may hold several)
    if r and r[0] == "Address":
        break
    if len(r) >= len(h):
        body.append(r)
if len(body) != len(lines):
    print(f"warning: {len(body)} SASS rows in the report vs {len(lines)} instructions in the cubin", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
ti = ts = 0
for r, l in zip(body, lines):
    inst, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    a = agg[l]
    a[0] += inst
    a[1] += s
    for c in stall_cols:
        v = int(r[ix[c]] or 0)
        if v:
            a[2][c[6:]] += v
    ti += inst
    ts += s
print(f"total warp instructions {ti}, samples {ts}")
for l, (i, s, st) in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    if 100 * i / ti >= minpct or 100 * s / ts >= minpct:
        top = ", ".join(f"{k} {100 * v / max(1, s):.0f}%" for k, v in st.most_common(3))
        print(f"{l[0] if l else '?':>18}:{l[1] if l else 0:<4} inst {100 * i / ti:5.1f}%  samples {100 * s / ts:5.1f}%   {top}")
