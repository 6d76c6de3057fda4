#!/usr/bin/env python
"""Microseconds per streamline point of path_2d_kernel on the device-resident field: the 18 635-point umass path
(one streamline, latency bound), the six golden starts in one launch, and 512 streamlines at once.
usage: path_timing.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402

import common  # noqa: E402
from epic_b200 import grids  # noqa: E402
from epic_b200.harmonic import Harmonic  # noqa: E402

u, locked, eps, stagger = common.case_input("umass")
h = Harmonic(u.copy(), locked.copy(), eps, stagger)
h.solve(process="gpu")
h.initialize_gpu()
starts = [(339, 184), (667, 152), (337, 24), (90, 197), (743, 242), (697, 202)]
n = u.size
for rep in range(2):
    t0 = time.perf_counter()
    r, p = h.compute_path(90.0, 197.0, 0.05, 0.5, int(n / 0.05), "gpu")
    dt = time.perf_counter() - t0
print("one streamline: %d points in %.2f ms = %.3f us/point (device path, incl. copies)" % (len(p), dt * 1e3, dt / len(p) * 1e6))
t0 = time.perf_counter()
r, pc = h.compute_path(90.0, 197.0, 0.05, 0.5, int(n / 0.05), "cpu")
dtc = time.perf_counter() - t0
print("same on the host export: %d points in %.2f ms = %.3f us/point; identical: %s" % (len(pc), dtc * 1e3, dtc / len(pc) * 1e6, np.array_equal(p, pc)))
t0 = time.perf_counter()
res = h.compute_paths_gpu(np.array(starts, np.float32), 0.05, 0.5, int(n / 0.05))
dt = time.perf_counter() - t0
pts = sum(len(q) for _, q in res)
print("six streamlines in one call: %d points in %.2f ms = %.3f us/point" % (pts, dt * 1e3, dt / pts * 1e6))
many = np.array(grids.free_cells(locked, 512, seed=3), np.float32)
t0 = time.perf_counter()
res = h.compute_paths_gpu(many, 0.2, 0.4, 1000000)
dt = time.perf_counter() - t0
pts = sum(len(q) for _, q in res)
print("512 streamlines in one call: %d points in %.2f ms = %.4f us/point" % (pts, dt * 1e3, dt / pts * 1e6))
h.uninitialize_gpu()
