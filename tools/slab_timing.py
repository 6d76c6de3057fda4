#!/usr/bin/env python
"""Pass time of ONE slab of a sharded grid on one GPU (what each rank of an N-GPU run executes per pass),
for several tile geometries.  usage: slab_timing.py [world] [size] [math ...]   CONFIGS=threads:tile_rows,..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from epic_b200 import grids  # noqa: E402
from epic_b200.field import Field  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
size = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
maths = sys.argv[3:] or ["strict", "fast"]
rows = size // world
row0 = rows * (world // 2)
ghost = 4
u, locked = grids.random_obstacles((size, size), 0.2, 64, seed=1234, row0=row0 - ghost, rows=rows + 2 * ghost)
for math in maths:
    for cfg in (os.environ.get("CONFIGS") or "0:0,256:32,256:40,256:48,256:56,256:64,512:64,512:80,512:96").split(","):
        nt, th = (int(x) for x in cfg.split(":"))
        for k, v in (("EPIC_THREADS", nt), ("EPIC_TILE_ROWS", th)):
            if v:
                os.environ[k] = str(v)
            else:
                os.environ.pop(k, None)
        f = Field((size, size), row0=row0, rows=rows, ghost=ghost, math=math, stream=torch.cuda.current_stream().cuda_stream)
        f.upload(u, locked, first=row0 - ghost, layers=rows + 2 * ghost)
        f.run(0, 200, False)
        f.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        f.run(200, 400, False)
        b.record()
        torch.cuda.synchronize()
        info = f.info()
        ms = a.elapsed_time(b) / 100
        print("slab %d rows of %d^2 %-6s threads %s tile_rows %d: %.4f ms/pass -> %.1f GCUPS per GPU (x%d = %.0f)" % (
            rows, size, math, nt or "auto", info["tile_rows"], ms, rows * size / 2 * 4 / (ms * 1e-3) / 1e9, world,
            world * rows * size / 2 * 4 / (ms * 1e-3) / 1e9), flush=True)
        f.close()
