#!/usr/bin/env python3
"""Static per-phase SASS instruction counts of a kernel in a cubin (no GPU needed): groups instructions by the
`// ---- ` phase comments of the .cu file via nvdisasm line info.  Straight-line unrolled code => static ~ dynamic.
usage: sass_static.py file.cubin kernel_substring source.cu"""
import collections, re, subprocess, sys

cubin, kname, src = sys.argv[1:4]
marks = [(1, "0 setup")]
for i, ln in enumerate(open(src), 1):
    m = re.match(r"\s*// ---- (.{0,28})", ln)
    if m and "kernel" not in ln:
        marks.append((i, f"{len(marks)} {m.group(1).strip()}"))
kstart = next(i for i, ln in enumerate(open(src), 1) if "__global__" in ln)
base = src.split("/")[-1]
dis = subprocess.run(["nvdisasm", "-c", "--print-line-info", cubin], capture_output=True, text=True).stdout
inside, cur, prev = False, None, "0 setup"
agg, byop = collections.Counter(), collections.defaultdict(collections.Counter)
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
    if not m:
        continue
    if cur and cur[0] == base and cur[1] >= kstart:
        for l0, nm in marks:
            if cur[1] >= l0:
                prev = nm
    t = m.group(1).split()
    op = t[1] if t[0].startswith("@") else t[0]
    agg[prev] += 1
    byop[prev][op] += 1
print("total", sum(agg.values()))
for p in sorted(agg, key=lambda x: int(x.split()[0])):
    print(f"{p:32s} {agg[p]:5d}   " + ", ".join(f"{k} {v}" for k, v in byop[p].most_common(12)))
