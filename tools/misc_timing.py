#!/usr/bin/env python
"""Wall-clock of the small-map and 3-D configurations of BASELINE.json on one B200:
time-to-epsilon of maps/maze.png and maps/umass.png through harmonic_complete_gpu, and the 3-D sweep rate."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import common  # noqa: E402
from epic_b200 import grids  # noqa: E402
from epic_b200.field import Field  # noqa: E402
from epic_b200.harmonic import Harmonic  # noqa: E402

for name in ("maze", "umass", "basic"):
    u, locked, eps, stagger = common.case_input(name)
    for rep in range(2):
        h = Harmonic(u.copy(), locked.copy(), eps, stagger)
        t0 = time.perf_counter()
        h.solve(process="gpu")
        dt = time.perf_counter() - t0
    n = u.size
    print("%-6s %s: complete_gpu %.3f s, %d iterations, delta %.6g, %.1f us/iteration, %.2f GCUPS" % (
        name, u.shape, dt, h.currentIteration, h.delta, dt / h.currentIteration * 1e6,
        n / 2 * h.currentIteration / dt / 1e9), flush=True)

for shape in ((512, 512, 512), (1024, 1024, 1024)):
    for math in ("strict", "fast"):
        u, locked = grids.random_obstacles(shape, 0.2, 64, seed=1234)
        f = Field(shape, math=math, stream=torch.cuda.current_stream().cuda_stream)
        f.upload(u, locked)
        del u, locked
        f.run(0, 300, False)
        f.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        f.run(300, 20, False)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        print("3-D %s %s: %.3f ms/sweep, %.1f GCUPS" % (shape, math, ms, np.prod(shape) / 2 / (ms * 1e-3) / 1e9), flush=True)
        f.close()
