#!/usr/bin/env python
"""Hottest source lines of ONE kernel of an ncu report (instructions executed + stall samples per CUDA line).
usage: scripts/ncu_lines.py <rep> <kernel regex> [top N]"""
import collections, csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "-k", "regex:" + kre, "-c", "1"],
                     capture_output=True, text=True).stdout
cur, h = None, None
inst, samp, text = collections.Counter(), collections.Counter(), {}
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur = r[1]; continue
    if r[0] == "Line No":
        h = r; continue
    if h and cur and r[0].isdigit() and cur.endswith((".cuh", ".cu")):
        d = dict(zip(h, r))
        key = (cur.split("/")[-1], int(r[0]))
        try:
            inst[key] += int(d.get("Instructions Executed") or 0)
            samp[key] += int(d.get("Warp Stall Sampling (All Samples)") or d.get("Warp Stall Sampling (All Cycles)") or 0)
        except ValueError:
            pass
        text[key] = r[1][:110]
ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
print(f"total warp instructions {ti}, stall samples {ts}")
for k, v in samp.most_common(top):
    print(f"{100*v/ts:5.1f}% samples {100*inst[k]/ti:5.1f}% inst  {k[0]}:{k[1]}  {text[k]}")
print("-- by instructions executed")
for k, v in inst.most_common(top):
    print(f"{100*v/ti:5.1f}% inst {100*samp[k]/ts:5.1f}% samples  {k[0]}:{k[1]}  {text[k]}")
