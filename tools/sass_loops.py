#!/usr/bin/env python
"""Backward-branch bodies (loops) of one kernel with their instruction mix.
usage: sass_loops.py lib.so kernel-substring [max_body]"""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
max_body = int(sys.argv[3]) if len(sys.argv) > 3 else 1400
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, ins = None, []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
print(pat, len(ins), "instructions")
for a, t in ins:
    if "BRA" in t:
        x = re.search(r"0x([0-9a-f]+)", t)
        if x and int(x.group(1), 16) < a:
            lo = int(x.group(1), 16)
            body = [tt for aa, tt in ins if lo <= aa <= a]
            c = collections.Counter(re.sub(r"^@!?\w+\s+", "", tt).split()[0].split(".")[0] for tt in body)
            if len(body) < max_body:
                print(hex(lo), hex(a), len(body), dict(c.most_common(18)))
