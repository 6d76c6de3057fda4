#!/usr/bin/env python
"""Generate tests/golden/ from the UNTOUCHED reference CPU code.

Runs only in the build container (needs /root/reference): the reference's CPU sources are
compiled where they lie into oracle/_ref/libepic_ref_cpu.so (oracle/Makefile) and driven through
their own C API (harmonic_complete_cpu, harmonic_update[_and_check]_cpu,
harmonic_compute_path_2d_cpu, harmonic_utilities_set_cells_2d_cpu).  Outputs:

  tests/golden/maps.npz     the two demo maps (maps/maze.png, maps/umass.png) and three of the
                            reference's test maps as uint8 images (inputs; the GPU box has no
                            /root/reference)
  tests/golden/golden.json  per case: iterations, delta (exact hex), sha1 of the field bytes,
                            probe values, path lengths / hashes / sampled points

Usage: python tests/golden/make_golden.py [case ...]     (no argument = all cases; slow ones ~6 min)
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from epic_b200 import grids  # noqa: E402
from oracle import oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def sha1(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def field_record(ref, probes=()):
    free = ref.locked == 0
    rec = {"iterations": int(ref.iteration), "delta_hex": float(ref.delta).hex(), "delta": float(ref.delta),
           "sha1_u": sha1(ref.u)}
    if free.any():
        rec["free_min"] = float(ref.u[free].min())
        rec["free_max"] = float(ref.u[free].max())
    rec["probes"] = [[list(map(int, p)), float(ref.u[tuple(p)]).hex()] for p in probes]
    return rec


def path_record(ref, x, y, step, cd, max_length):
    r, p = ref.path(float(x), float(y), step, cd, max_length)
    rec = {"start": [float(x), float(y)], "step": step, "cd": cd, "max_length": int(max_length), "ret": int(r),
           "k": int(len(p))}
    if r == 0:
        cells = np.floor(p + np.float32(0.5)).astype(np.int64)
        keep = np.ones(len(cells), bool)
        keep[1:] = (cells[1:] != cells[:-1]).any(axis=1)
        rec.update({"sha1_path": sha1(p), "first": [float(v).hex() for v in p[1]],
                    "last": [float(v).hex() for v in p[-1]], "distinct_cells": int(keep.sum()),
                    "sha1_cells": sha1(cells[keep]),
                    "every_64th": [[float(a).hex(), float(b).hex()] for a, b in p[::64][:64]]})
    return rec


def checkpoints(u, locked, ks, eps=1e-3, stagger=100):
    """sha1 of the field after exactly k iterations of the complete() schedule, k in ks."""
    ref = orc.Reference(u.copy(), locked.copy(), eps, stagger)
    out, done = {}, 0
    for k in sorted(ks):
        ref.run_iterations(k - done)
        done = k
        out[str(k)] = {"sha1_u": sha1(ref.u), "delta_hex": float(ref.delta).hex()}
    return out


def load_maps():
    import cv2
    names = {"maze": "maps/maze.png", "umass": "maps/umass.png",
             "basic": "libepic/tests/maps/basic.png", "trivial": "libepic/tests/maps/trivial.png",
             "c_space": "libepic/tests/maps/c_space.png"}
    return {k: cv2.imread(os.path.join(REF, v), cv2.IMREAD_GRAYSCALE) for k, v in names.items()}


def case_map(name, images, starts, full=True):
    u, locked = grids.grid_from_image(images[name])
    rec = {"input": "maps.npz[%s] via grids.grid_from_image" % name, "shape": list(u.shape), "epsilon": 1e-3,
           "stagger": 100, "checkpoints": checkpoints(u, locked, [1, 2, 3, 10, 100, 101, 250])}
    if full:
        ref = orc.Reference(u.copy(), locked.copy(), 1e-3, 100)
        t0 = time.time()
        ret = ref.complete()
        rec["complete"] = dict(field_record(ref), ret=int(ret), seconds=round(time.time() - t0, 1))
        n = u.size
        rec["paths"] = [path_record(ref, x, y, 0.05, 0.5, int(n / 0.05)) for x, y in starts]
        rec["paths"] += [path_record(ref, x, y, 0.2, 0.4, 1000000) for x, y in starts]
        rec["potentials"] = potential_records(ref, locked, seed=len(name))
    return rec


def potential_records(ref, locked, seed, count=40):
    """Bilinear potentials and central-difference gradients (harmonic_path_cpu.cpp:41-118) at seeded sub-cell
    positions: free cells with fractional offsets (some next to obstacles, where a tap or the gradient fails),
    plus positions on obstacles and at the border."""
    rng = np.random.RandomState(1000 + seed)
    pts = []
    for x, y in grids.free_cells(locked, count, seed=seed + 3):
        pts.append((np.float32(x + rng.uniform(-0.49, 0.49)), np.float32(y + rng.uniform(-0.49, 0.49))))
    ys, xs = np.nonzero(locked == 1)
    for i in rng.choice(len(ys), 6, replace=False):
        pts.append((np.float32(xs[i] + 0.25), np.float32(ys[i] - 0.25)))
    pts += [(np.float32(0.2), np.float32(locked.shape[0] / 2)), (np.float32(locked.shape[1] - 0.6), np.float32(1.5))]
    out = []
    for x, y in pts:
        x, y = float(x), float(y)
        r, v = ref.potential(x, y)
        rg, gx, gy = ref.gradient(x, y, 0.5)
        out.append({"xy": [x, y], "ret": int(r), "value_hex": float(v).hex(), "grad_ret": int(rg),
                    "grad_hex": [float(gx).hex(), float(gy).hex()]})
    return out


def case_box64():
    u = np.full((64, 64), -1e6, np.float32)
    locked = np.zeros((64, 64), np.uint32)
    locked[0, :] = locked[-1, :] = locked[:, 0] = locked[:, -1] = 1
    u[10, 10], locked[10, 10] = 0.0, 1
    ref = orc.Reference(u.copy(), locked.copy(), 1e-3, 100)
    ret = ref.complete()
    rec = {"input": "64x64, border locked at -1e6, goal cell (row 10, col 10)", "epsilon": 1e-3, "stagger": 100,
           "checkpoints": checkpoints(u, locked, [1, 2, 7, 100, 101]),
           "complete": dict(field_record(ref, probes=[(10, 11), (9, 10), (50, 50), (62, 62)]), ret=int(ret))}
    rec["paths"] = [path_record(ref, 50.0, 50.0, 0.05, 0.5, int(64 * 64 / 0.05)),
                    path_record(ref, 30.0, 55.0, 0.2, 0.4, 1000000),
                    path_record(ref, 0.0, 5.0, 0.05, 0.5, 1000),      # starts on an obstacle
                    path_record(ref, 10.0, 10.0, 0.05, 0.5, 1000),    # starts on the goal
                    path_record(ref, 50.0, 50.0, 0.05, 0.5, 40)]      # truncated by max_length
    pts = [(50.0, 50.0), (10.4, 10.6), (1.2, 1.3), (30.5, 30.5), (0.2, 30.0), (62.6, 62.6)]
    rec["potentials"] = []
    for x, y in pts:
        r, v = ref.potential(x, y)
        rg, gx, gy = ref.gradient(x, y, 0.5)
        rec["potentials"].append({"xy": [x, y], "ret": int(r), "value_hex": float(v).hex(), "grad_ret": int(rg),
                                  "grad_hex": [float(gx).hex(), float(gy).hex()]})
    return rec


def case_synthetic(shape, p, goals, seed, eps=1e-3, stagger=100, ks=(1, 2, 3, 50, 101)):
    u, locked = grids.random_obstacles(shape, p, goals, seed)
    ref = orc.Reference(u.copy(), locked.copy(), eps, stagger)
    ret = ref.complete()
    return {"input": "grids.random_obstacles(%r, p=%r, goals=%d, seed=%d)" % (tuple(shape), p, goals, seed),
            "epsilon": eps, "stagger": stagger, "sha1_u0": sha1(u), "sha1_locked": sha1(locked),
            "checkpoints": checkpoints(u, locked, ks, eps, stagger), "complete": dict(field_record(ref), ret=int(ret))}


def case_proc_maze():
    u, locked = grids.procedural_maze((200, 300), corridor=8, wall=2, goals=3, seed=99)
    ref = orc.Reference(u.copy(), locked.copy(), 1e-3, 100)
    ret = ref.complete()
    rec = {"input": "grids.procedural_maze((200, 300), corridor=8, wall=2, goals=3, seed=99)", "epsilon": 1e-3,
           "stagger": 100, "sha1_u0": sha1(u), "sha1_locked": sha1(locked),
           "checkpoints": checkpoints(u, locked, [1, 2, 64, 101]), "complete": dict(field_record(ref), ret=int(ret))}
    rec["paths"] = [path_record(ref, x, y, 0.05, 0.5, int(u.size / 0.05)) for x, y in grids.free_cells(locked, 6)]
    return rec


def case_set_cells():
    """Anytime-node style: iterate, edit cells (goal/obstacle/free, some out of range), iterate."""
    u, locked = grids.random_obstacles((96, 130), 0.15, 3, 5)
    ref = orc.Reference(u.copy(), locked.copy(), 1e-3, 50)
    ref.run_iterations(120)
    rng = np.random.RandomState(11)
    k = 60
    v = np.stack([rng.randint(0, 140, size=k), rng.randint(0, 100, size=k)], axis=1).astype(np.uint32)
    types = rng.randint(0, 4, size=k).astype(np.uint32)   # 3 = invalid type, some coordinates out of range
    ret = ref.set_cells(v, types)
    mid = sha1(ref.u), sha1(ref.locked)
    ref.run_iterations(130)
    return {"input": "grids.random_obstacles((96,130),0.15,3,5); 120 its; set_cells(RandomState(11)); 130 its",
            "stagger": 50, "v": v.tolist(), "types": types.tolist(), "ret": int(ret), "sha1_u_after_edit": mid[0],
            "sha1_locked_after_edit": mid[1], "final": field_record(ref), "sha1_locked_final": sha1(ref.locked)}


def main():
    os.makedirs(GOLD, exist_ok=True)
    assert orc.have_ref(), "needs /root/reference (build container)"
    images = load_maps()
    np.savez_compressed(os.path.join(GOLD, "maps.npz"), **images)
    path = os.path.join(GOLD, "golden.json")
    gold = json.load(open(path)) if os.path.exists(path) else {}
    umass_starts = [(339, 184), (667, 152), (337, 24), (90, 197), (743, 242), (697, 202)]
    maze_u, maze_l = grids.grid_from_image(images["maze"])
    cases = {
        "box64": case_box64,
        "random256": lambda: case_synthetic((256, 256), 0.2, 8, 1234),
        "random_ragged": lambda: case_synthetic((131, 77), 0.1, 2, 3, ks=(1, 2, 3, 9)),
        "random48x3": lambda: case_synthetic((48, 48, 48), 0.2, 8, 1234),
        "random3d_ragged": lambda: case_synthetic((21, 34, 45), 0.1, 2, 8, ks=(1, 2, 5)),
        "proc_maze": case_proc_maze,
        "set_cells": case_set_cells,
        "basic": lambda: case_map("basic", images, grids.free_cells(grids.grid_from_image(images["basic"])[1], 4)),
        "umass": lambda: case_map("umass", images, umass_starts),
        "maze": lambda: case_map("maze", images, grids.free_cells(maze_l, 6)),
    }
    todo = sys.argv[1:] or list(cases)
    for name in todo:
        t0 = time.time()
        gold[name] = cases[name]()
        json.dump(gold, open(path, "w"), indent=1, sort_keys=True)
        print("%-16s %.1fs" % (name, time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
