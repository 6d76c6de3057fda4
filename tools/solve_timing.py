#!/usr/bin/env python
"""Wall-clock of harmonic_complete_gpu (upload, solve to epsilon, download) on the demo maps, several
repetitions each: shows the host-side launch cost / jitter that solver periods replayed from CUDA graphs remove.
usage: solve_timing.py [reps]     (EPIC_SWEEPS_PER_PASS / EPIC_TILE_ROWS from the environment)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import common  # noqa: E402
from epic_b200.harmonic import Harmonic  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for name in ("maze", "umass", "basic"):
    u, locked, eps, stagger = common.case_input(name)
    times = []
    for rep in range(reps + 1):
        h = Harmonic(u.copy(), locked.copy(), eps, stagger)
        t0 = time.perf_counter()
        h.solve(process="gpu")
        times.append(time.perf_counter() - t0)
    times = times[1:]   # the first call pays context creation / module load
    print("%-6s %s: %d iterations, complete_gpu min %.3f s  median %.3f s  max %.3f s  (%.2f us/iteration at best)" % (
        name, u.shape, h.currentIteration, min(times), sorted(times)[len(times) // 2], max(times),
        min(times) / h.currentIteration * 1e6), flush=True)
