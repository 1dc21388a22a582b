#!/usr/bin/env python
"""Per-kernel SASS mnemonic histogram of libtmkcl.so (evidence for profiles/: which kernels carry TMA bulk copies
(UBLKCP), mbarrier waits (SYNCS), FP64 (DFMA/DADD/DMUL), and how large each kernel is).
usage: scripts/sass_summary.py [path to .so] > profiles/<round>_sass_summary.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "tmkit_b200/libtmkcl.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, hist, arch = None, collections.OrderedDict(), collections.Counter()
for line in txt.splitlines():
    m = re.search(r"arch = (\S+)", line)
    if m:
        arch[m.group(1)] += 1
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
dem = subprocess.run(["cu++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
print("cubins:", dict(arch))
cols = ["UBLKCP", "SYNCS", "UTMALDG", "LDG", "STG", "LDS", "STS", "LDC", "ULDC", "FFMA", "FMUL", "FADD", "DFMA", "MUFU", "VOTE", "ATOMS", "ATOMG", "RED", "BAR", "LDL", "STL"]
print(f"{'kernel':86s} {'insts':>7s} " + " ".join(f"{c:>6s}" for c in cols))
tot = collections.Counter()
for (k, h), d in zip(hist.items(), dem):
    name = re.sub(r"^void ", "", d).replace("tmk::", "")
    name = re.sub(r"\(.*", "", name)[:86]
    print(f"{name:86s} {sum(h.values()):7d} " + " ".join(f"{h.get(c, 0):6d}" for c in cols))
    tot.update(h)
print(f"{'TOTAL':86s} {sum(tot.values()):7d} " + " ".join(f"{tot.get(c, 0):6d}" for c in cols))
