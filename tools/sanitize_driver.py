#!/usr/bin/env python
"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
2-D sweeps plain, with static-tile tracking (a solve), sharded over two slabs behind the ABI (peer stores, in-kernel
flags, the all-reduce decision kernels), 3-D sweeps, set-cells, map ingest, streamlines.  Results are compared with
the oracle so that a sanitizer-clean run is also a correct one.
usage: sanitize_driver.py [which ...]    which in: plain solve sharded sharded3d d3 paths edits (default: all)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import common  # noqa: E402
from epic_b200 import grids  # noqa: E402
from epic_b200.harmonic import Harmonic  # noqa: E402
from oracle import oracle as orc  # noqa: E402

os.environ["EPIC_MIN_SLAB_CELLS"] = "0"     # EPIC_DEVICES lists are taken literally, however small the grid
which = sys.argv[1:] or ["plain", "solve", "sharded", "sharded3d", "d3", "paths", "edits"]


def run_iterations(shape, n, devices=None, p=0.2, goals=3, seed=9):
    if devices:
        os.environ["EPIC_DEVICES"] = devices
    else:
        os.environ.pop("EPIC_DEVICES", None)
    u, locked = grids.random_obstacles(shape, p, goals, seed=seed)
    s = common.LibepicSolver(u.copy(), locked.copy(), 1e-3, 10, "gpu")
    o = orc.Oracle(u.copy(), locked.copy(), 1e-3, 10)
    s.run_iterations(n)
    o.run_iterations(n)
    ok = np.array_equal(s.u, o.u) and s.delta == o.delta
    s.close()
    return ok


def solve(shape, devices=None, seed=4):
    if devices:
        os.environ["EPIC_DEVICES"] = devices
    else:
        os.environ.pop("EPIC_DEVICES", None)
    u, locked = grids.random_obstacles(shape, 0.15, 2, seed=seed)
    h = Harmonic(u.copy(), locked.copy(), 1e-2, 10)
    h.solve(process="gpu")
    o = orc.Oracle(u.copy(), locked.copy(), 1e-2, 10)
    o.complete()
    return h.currentIteration == o.iteration and np.array_equal(h.field, o.u)


results = {}
if "plain" in which:
    results["plain 2-D sweeps (70 x 300, 23 iterations)"] = run_iterations((70, 300), 23)
if "solve" in which:
    results["2-D solve with static-tile tracking (60 x 280)"] = solve((60, 280))
if "sharded" in which:
    results["2-D sharded behind the ABI, 2 slabs (peer stores, flags), 23 iterations"] = run_iterations((96, 300), 23, "0,0")
    results["2-D sharded solve (decision all-reduce), 3 slabs"] = solve((120, 280), "0,0,0")
if "d3" in which:
    results["3-D sweeps (20 x 24 x 140, 9 iterations)"] = run_iterations((20, 24, 140), 9)
if "sharded3d" in which:
    results["3-D sharded behind the ABI, 2 slabs, 9 iterations"] = run_iterations((24, 24, 140), 9, "0,0")
os.environ.pop("EPIC_DEVICES", None)
if "paths" in which:
    u, locked, eps, stagger = common.case_input("box64")
    s = common.LibepicSolver(u.copy(), locked.copy(), 1e-2, stagger, "gpu", paths_on="gpu")
    o = orc.Oracle(u.copy(), locked.copy(), 1e-2, stagger)
    s.complete()
    o.complete()
    ok = True
    for start in ((50.0, 50.0), (12.0, 55.5)):
        rp, p = s.path(*start, 0.2, 0.4, 5000)
        ro, po = o.path(*start, 0.2, 0.4, 5000)
        ok = ok and rp == ro and np.array_equal(p, po)
    ok = ok and s.potential(30.3, 31.7) == o.potential(30.3, 31.7) and s.gradient(30.3, 31.7, 0.5) == o.gradient(30.3, 31.7, 0.5)
    s.close()
    results["streamlines / potential / gradient on the device-resident field"] = ok
if "edits" in which:
    u, locked = grids.random_obstacles((64, 200), 0.15, 3, seed=5)
    s = common.LibepicSolver(u.copy(), locked.copy(), 1e-3, 10, "gpu")
    o = orc.Oracle(u.copy(), locked.copy(), 1e-3, 10)
    s.run_iterations(6)
    o.run_iterations(6)
    rng = np.random.RandomState(2)
    v = np.stack([rng.randint(0, 210, 40), rng.randint(0, 70, 40)], 1).astype(np.uint32)
    t = rng.randint(0, 4, 40).astype(np.uint32)
    s.set_cells(v, t)
    o.set_cells(v, t)
    occ = np.where(o.locked == 1, 100, 0).astype(np.int8)
    occ[20:30, 50:90] = 100
    s.h.set_occupancy_grid(occ, "cpu")
    s.h.set_occupancy_grid(occ, "gpu")
    s.run_iterations(5)
    s.h.reset_free_cells("cpu")
    s.h.reset_free_cells("gpu")
    s.run_iterations(5)
    got = s.u.copy()
    hc = Harmonic(u.copy(), locked.copy(), 1e-3, 10)   # the same sequence on the library's CPU exports
    hc.run_iterations(6, "cpu")
    hc.set_cells(v, t, "cpu")
    hc.set_occupancy_grid(occ, "cpu")
    hc.run_iterations(5, "cpu")
    hc.reset_free_cells("cpu")
    hc.run_iterations(5, "cpu")
    results["set-cells / occupancy ingest / reset-free-cells"] = np.array_equal(got, hc.field)
    s.close()
bad = [k for k, v in results.items() if not v]
for k, v in results.items():
    print("%-75s %s" % (k, "ok" if v else "MISMATCH"))
sys.exit(1 if bad else 0)
