#!/usr/bin/env python3
"""profiles/traffic.json (DRAM bytes per launch of the kernels bench.py reports a roofline for) from the summary of an
`ncu --set full` capture of `bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --skip-sw` (scripts/r2_profile.sh).
Every entry carries the sha1 of the kernel's source files at capture time: bench.py drops an entry whose sources have changed.
usage: make_traffic.py ncu_full_summary_TAG.csv TAG"""
import csv, hashlib, json, pathlib, sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
rows = list(csv.reader(open(sys.argv[1])))
tag = sys.argv[2]
h = rows[0]
ki = h.index("kernel")
ri = next(i for i, c in enumerate(h) if c.startswith("dram__bytes_read.sum"))
wi = next(i for i, c in enumerate(h) if c.startswith("dram__bytes_write.sum"))
assert "Gbyte" in h[ri] and "Gbyte" in h[wi], (h[ri], h[wi])
last = {}
for r in rows[1:]:
    last[r[ki]] = (float(r[ri]) + float(r[wi])) * 1e9          # the last launch of each kernel (warm caches, final buffers)


def total(*names):
    return sum(v for k, v in last.items() if any(n in k for n in names))


def sha(*files):
    return {f: hashlib.sha1((ROOT / "gams_b200" / "csrc" / f).read_bytes()).hexdigest() for f in files}


out = {
    "source": f"profiles/{tag}_ncu_full_summary.csv (ncu --set full --clock-control none, bench.py --steps 2 --warmup 1 --skip-e2e "
              "--skip-cpu --skip-sw; last launch of each kernel)",
    "gc_wave_fused_kernel": {"dram_bytes_per_launch": total("gc_wave_fused_kernel"), "source_sha1": sha("gc_wave_fused.cu", "gc_common.cuh")},
    "gc_window_kernel": {"dram_bytes_per_launch": total("gc_window_kernel"), "source_sha1": sha("gc_window.cu", "gc_common.cuh")},
    "zscore_kernel": {"dram_bytes_per_launch": total("zscore_indep_kernel"), "source_sha1": sha("zscore.cu", "zscore_common.cuh")},
    "count_overlaps:grouped_by_ctg": {"dram_bytes_per_launch": total("seg_query_bounds_kernel", "seg_items_kernel", "count_overlaps_staged_kernel"),
                                      "note": "seg_query_bounds_kernel + seg_items_kernel + count_overlaps_staged_kernel",
                                      "source_sha1": sha("overlap.cu", "overlap_index.cuh")},
    "count_overlaps:arbitrary_order": {"dram_bytes_per_launch": total("part_sort_kernel", "part_lookup_kernel", "part_unpermute_kernel"),
                                       "note": "part_sort_kernel + part_lookup_kernel + part_unpermute_kernel",
                                       "source_sha1": sha("overlap_part.cu", "overlap_index.cuh")},
    "count_windows": {"dram_bytes_per_launch": total("cw_tile_kernel"), "note": "cw_tile_kernel", "source_sha1": sha("count_windows.cu")},
}
(ROOT / "profiles" / "traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps({k: v["dram_bytes_per_launch"] for k, v in out.items() if isinstance(v, dict)}, indent=1))
