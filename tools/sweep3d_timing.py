#!/usr/bin/env python
"""Device-resident 3-D sweep throughput (CUDA events around N passes) for a few tile heights.
usage: sweep3d_timing.py [size] [math ...]      CONFIGS=tile_rows,... (0 = default)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from epic_b200 import grids  # noqa: E402
from epic_b200.field import Field  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
maths = sys.argv[2:] or ["strict", "fast"]
shape = (size, size, size)
u, locked = grids.random_obstacles(shape, 0.2, 64, seed=1234)
for math in maths:
    for th in (os.environ.get("CONFIGS") or "0").split(","):
        if int(th):
            os.environ["EPIC_TILE_ROWS"] = th
        else:
            os.environ.pop("EPIC_TILE_ROWS", None)
        f = Field(shape, math=math, stream=torch.cuda.current_stream().cuda_stream)
        f.upload(u, locked)
        f.run(0, 300, False)
        f.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        f.run(300, 40, False)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 40
        print("3-D %s %-6s tile_rows %d: %.3f ms/sweep, %.1f GCUPS" % (
            shape, math, f.info()["tile_rows"], ms, np.prod(shape) / 2 / (ms * 1e-3) / 1e9), flush=True)
        f.close()
