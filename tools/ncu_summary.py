#!/usr/bin/env python
"""Summarise an .ncu-rep (from `ncu --set full --import-source on`) into a small text file for profiles/.
usage: ncu_summary.py report.ncu-rep out.md [title]"""
import collections
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max"]
lines = ["# %s" % title, "", "source: `%s` (ncu --set full --clock-control none --import-source on)" % rep, "",
         "| metric | unit | " + " | ".join("launch %d" % i for i in range(len(data))) + " |", "|---|---|" + "---|" * len(data)]
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        lines.append("| %s | %s | %s |" % (k, units[i], " | ".join(r[i][:90] for r in data)))
lines += ["", "warp stall reasons (cycles per issued instruction):", ""]
for i, h in enumerate(hdr):
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        name = h.split("issue_stalled_")[1].split("_per_issue")[0]
        v = float(data[0][i])
        if v >= 0.03:
            lines.append("* %s: %.2f" % (name, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(srows) if r and r[0] == "Address")
sh = srows[h]
ia, isrc = sh.index("Instructions Executed"), sh.index("Source")
body = [r for r in srows[h + 1:] if len(r) > ia and r[0].startswith("0x")]
first = []
seen = set()
for r in body:          # the source page lists every captured launch back to back: keep the first
    if r[0] in seen:
        break
    seen.add(r[0])
    first.append(r)
tot = sum(int(r[ia]) for r in first)
mix = collections.Counter()
for r in first:
    t = r[isrc].split()
    op = t[1] if t[0].startswith("@") else t[0]
    mix[op.split(".")[0]] += int(r[ia])
lines += ["", "executed warp-instruction mix (first captured launch, %d SASS instructions, %.3g executed):" % (len(first), tot), ""]
for k, v in mix.most_common(22):
    lines.append("* %s: %.1f %%" % (k, 100.0 * v / tot))
tma = [r[isrc].strip() for r in first if "UTMALDG" in r[isrc] or "UTMAPF" in r[isrc] or "UTMASTG" in r[isrc]]
lines += ["", "TMA instructions in the SASS: " + (", ".join(sorted(set(t.split()[0] if not t.startswith("@") else t.split()[1] for t in tma))) or "none")]
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
