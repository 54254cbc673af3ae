#!/usr/bin/env python3
"""Key metrics of the first kernel in an ncu report. usage: ncu_key.py report.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h, u = r[0], r[1]
KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_l1tex2xbar_write_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed")
for row in r[2:]:
    print("==", row[h.index("Kernel Name")][:70])
    for a, b, c in zip(h, u, row):
        if a in KEYS or (a.startswith("sm__inst_executed_pipe") and a.endswith("avg.pct_of_peak_sustained_active") and c and float(c) > 1):
            print(f"  {a} [{b}] {c}")
