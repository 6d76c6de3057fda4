#!/usr/bin/env python
"""Small driver for ncu: a 16384^2 random-obstacle field, a few hundred warm sweeps so the field is in
its steady state (all three non-max neighbours contribute), then a handful of passes to capture.
usage: profile_sweep.py [strict|fast] [size | AxBxC] [warm_sweeps] [passes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from epic_b200 import grids  # noqa: E402
from epic_b200.field import Field  # noqa: E402

math = sys.argv[1] if len(sys.argv) > 1 else "strict"
size = sys.argv[2] if len(sys.argv) > 2 else "16384"
shape = tuple(int(x) for x in size.split("x")) if "x" in size else (int(size), int(size))
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 400
passes = int(sys.argv[4]) if len(sys.argv) > 4 else 6
u, locked = grids.random_obstacles(shape, 0.2, 64, seed=1234)
f = Field(shape, math=math)
f.upload(u, locked)
f.run(0, warm, False)
f.sync()
T = f.info()["sweeps_per_pass"]
for i in range(passes):
    f.run(warm + i * T, T, check_last=(i == passes - 1))
print("delta", f.read_delta(), f.info())
