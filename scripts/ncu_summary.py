#!/usr/bin/env python
"""Summarise an ncu report (raw + source pages) the way profiles/README.md quotes it.
usage: scripts/ncu_summary.py gpurun_out/<name>.ncu-rep [launch index]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
d = dict(zip(hdr, rows[2 + idx]))
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in want:
    if k in d:
        print(f"{k:75s} {d[k][:90]} {units[hdr.index(k)]}")
for k in hdr:
    if "average_warps_issue_stalled" in k and "not_issued" not in k:
        try:
            v = float(d[k])
        except ValueError:
            continue
        if v > 0.25:
            print(f"{k:75s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur = None
h = None
agg = collections.Counter()
lines = {}
tot = 0
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
        continue
    if r[0] == "Line No":
        h = r
        continue
    if h and cur and len(r) >= 8 and cur.endswith((".cuh", ".cu")) and r[0].isdigit():
        dd = dict(zip(h, r))
        try:
            n = int(dd["Instructions Executed"] or 0)
        except ValueError:
            continue
        key = (cur.split("/")[-1], int(r[0]))
        agg[key] += n
        lines[key] = r[1][:90]
        tot += n
if tot:
    byfile = collections.Counter()
    for (f, l), v in agg.items():
        byfile[f] += v
    print("instruction share by file:", {f: round(100 * v / tot, 1) for f, v in byfile.items()})

    def rng(f, a, b):
        return 100 * sum(v for (ff, l), v in agg.items() if ff == f and a <= l <= b) / tot
    import re
    marks = {}
    for f in byfile:
        try:
            txt = open(f"tmkit_b200/csrc/{f}").read().split("\n")
        except OSError:
            continue
        ms = [(i + 1, t) for i, t in enumerate(txt) if t.lstrip().startswith("/* ----")]
        ms.append((len(txt) + 1, ""))
        for (a, t), (b, _) in zip(ms, ms[1:]):
            sh = rng(f, a, b - 1)
            if sh > 0.3:
                print(f"  {sh:5.1f}%  {f}:{a}-{b-1}  {t.strip()[:80]}")
    for k, v in agg.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 12):
        print(f"{100*v/tot:5.2f}%  {k[0]}:{k[1]}  {lines[k]}")
