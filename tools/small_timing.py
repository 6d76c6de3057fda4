#!/usr/bin/env python
"""Pass time on the small demo maps for several tile geometries (the launch-latency-bound regime).
usage: small_timing.py [maze|umass|basic ...]   CONFIGS=threads:tile_rows,... (0:0 = the cost model's choice)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

import common  # noqa: E402
from epic_b200.field import Field  # noqa: E402

names = sys.argv[1:] or ["maze", "umass"]
cfgs = (os.environ.get("CONFIGS") or "0:0,256:16,256:24,256:32,256:48,512:64,512:96").split(",")
for name in names:
    u, locked, eps, stagger = common.case_input(name)
    for math in ("strict", "fast"):
        for cfg in cfgs:
            nt, th = (int(x) for x in cfg.split(":"))
            for k, v in (("EPIC_THREADS", nt), ("EPIC_TILE_ROWS", th)):
                if v:
                    os.environ[k] = str(v)
                else:
                    os.environ.pop(k, None)
            f = Field(u.shape, math=math, stream=torch.cuda.current_stream().cuda_stream)
            f.upload(u, locked)
            f.run(0, 2000, False)
            f.sync()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            f.run(2000, 4000, False)
            b.record()
            torch.cuda.synchronize()
            info = f.info()
            us = a.elapsed_time(b) * 1e3 / (4000 / info["sweeps_per_pass"])
            print("%-6s %s %-6s threads %s tile_rows %d: %.2f us/pass, %.2f GCUPS" % (
                name, u.shape, math, nt or "auto", info["tile_rows"], us,
                u.size / 2 * info["sweeps_per_pass"] / (us * 1e-6) / 1e9), flush=True)
            f.close()
