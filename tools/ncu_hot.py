#!/usr/bin/env python
"""Where a kernel's issue slots go: executed warp instructions and stall samples per SASS region of an
.ncu-rep captured with --import-source on.  Regions are split at branch targets / backward branches.
usage: ncu_hot.py report.ncu-rep [min_share_percent]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
recs = []
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None:
            break  # first captured launch only
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        recs.append(r)
iA, iS, iN, iX = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(recs[0][iA], 16)
ins = [(int(r[iA], 16) - base, r[iS].strip(), int(r[iN] or 0), int(r[iX] or 0)) for r in recs]
total_x = sum(i[3] for i in ins)
total_s = sum(i[2] for i in ins)
# split into regions where the executed count changes by more than 2 %
regions, start = [], 0
for k in range(1, len(ins) + 1):
    if k == len(ins) or abs(ins[k][3] - ins[k - 1][3]) > 0.02 * max(ins[k][3], ins[k - 1][3], 1):
        regions.append((start, k))
        start = k
print("total executed %.3e warp instructions, %d samples, %d SASS instructions" % (total_x, total_s, len(ins)))
# merge small adjacent regions into chunks of >= 24 instructions for readability
merged = []
for a, b in regions:
    if merged and (merged[-1][1] - merged[-1][0] < 24 or b - a < 24) and False:
        merged[-1] = (merged[-1][0], b)
    else:
        merged.append((a, b))
acc = []
for a, b in merged:
    x = sum(i[3] for i in ins[a:b])
    s = sum(i[2] for i in ins[a:b])
    acc.append((a, b, x, s))
# coalesce: print big regions individually, group the rest between them
grp = None
def flush(g):
    if g and 100.0 * g[2] / total_x >= min_share:
        print("0x%05x-0x%05x %5d instr  exec/instr %.3e  share %5.1f %%  samples %5.1f %%   [%s ... ]" % (
            ins[g[0]][0], ins[g[1] - 1][0], g[1] - g[0], g[2] / (g[1] - g[0]), 100.0 * g[2] / total_x, 100.0 * g[3] / max(total_s, 1),
            ins[g[0]][1][:40]))
for a, b, x, s in acc:
    if b - a >= 16:
        flush(grp)
        grp = None
        flush((a, b, x, s))
    else:
        grp = (grp[0], b, grp[2] + x, grp[3] + s) if grp else (a, b, x, s)
flush(grp)
