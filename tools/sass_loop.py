#!/usr/bin/env python
"""Instruction mix of one kernel's SASS and of its hottest loop (the largest backward-branch body).
usage: sass_loop.py lib.so kernel-substring"""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name:
        continue
    print("==", name, len(ins), "instructions")
    loops = []
    for addr, text in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?(\w+)\)?", text)
        if "BRA" in text:
            t = re.search(r"0x([0-9a-f]+)", text)
            if t and int(t.group(1), 16) < addr:
                loops.append((addr - int(t.group(1), 16), int(t.group(1), 16), addr))
    loops.sort(reverse=True)
    for size, lo, hi in loops[:4]:
        body = [t for a, t in ins if lo <= a <= hi]
        c = collections.Counter(re.sub(r"^@!?\w+\s+", "", t).split()[0].split(".")[0] for t in body)
        print(" loop 0x%x-0x%x: %d instructions" % (lo, hi, len(body)))
        print("   ", ", ".join("%s %d" % kv for kv in c.most_common(24)))
