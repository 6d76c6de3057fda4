"""Build and run tests/native/replay_callers.cpp (the ROS-free replay of the nav_core plugin and of the
anytime node) and the inputs of its two scenarios.  Shared by tests/test_replay.py and
tests/golden/make_replay_golden.py."""
import os
import subprocess

import numpy as np

import common
from epic_b200 import grids

SRC = os.path.join(common.ROOT, "tests", "native", "replay_callers.cpp")
OURS_INC = os.path.join(common.ROOT, "include")
OURS_LIBDIR = os.path.join(common.ROOT, "epic_b200", "lib")
REF_INC = "/root/reference/libepic/include"
REF_LIBDIR = os.path.join(common.ROOT, "oracle", "_ref")
GOLDEN = os.path.join(common.ROOT, "tests", "golden", "replay_callers.json")

# origin / resolution compiled into the harness
OX, OY, RES = -12.5, 3.25, 0.05


def build(out, include=OURS_INC, libdir=OURS_LIBDIR, lib="epic", cpu_only=False, dense_ingest=False):
    """g++ with the reference Makefile's arithmetic (no FMA contraction, no fast-math)."""
    cmd = ["g++", "-std=c++11", "-O2", "-ffp-contract=off", "-I", include, SRC, "-o", out,
           "-L", libdir, "-l" + lib, "-Wl,-rpath," + libdir, "-lm"]
    if cpu_only:
        cmd.insert(1, "-DREPLAY_CPU_ONLY")
    if dense_ingest:
        cmd.insert(1, "-DREPLAY_DENSE_INGEST")
    subprocess.run(cmd, check=True)
    return out


def write_map(path, cells):
    cells = np.ascontiguousarray(cells)
    with open(path, "wb") as f:
        f.write(np.array(cells.shape, np.uint32).tobytes())
        f.write(cells.astype(np.uint8, copy=False).tobytes() if cells.dtype != np.int8 else cells.tobytes())


def world(cell):
    """World coordinate of the centre-ish of a cell along either axis (x uses OX, y uses OY at the call site)."""
    return cell * RES


def _far_apart(free):
    """A goal cell and a start cell a good way apart, both in free space."""
    return free[len(free) // 9], free[-len(free) // 6]


def plan_case(tmp):
    """The plugin on a seeded 96 x 128 procedural maze as a costmap: lethal cells (254) on the walls,
    inflation-like costs below the plugin's threshold (250) in the corridors."""
    u, locked = grids.procedural_maze((96, 128), corridor=7, wall=2, goals=0, seed=21)
    rng = np.random.RandomState(5)
    cost = np.where(locked == 1, 254, rng.randint(0, 250, size=locked.shape)).astype(np.uint8)
    path = os.path.join(tmp, "plan_map.bin")
    write_map(path, cost)
    (gy, gx), (sy, sx) = _far_apart(np.argwhere(locked[1:-1, 1:-1] == 0) + 1)
    return ["plan", path, None, str(int(gx)), str(int(gy)), "%.4f" % (OX + (sx + 0.25) * RES), "%.4f" % (OY + (sy + 0.25) * RES)]


def plan_umass_case(tmp):
    """BASELINE.json config 2(ii): the plugin on maps/umass.png as map_server + costmap_2d hand it over -- black
    pixels are lethal (254), everything else (150 = unknown, 255 = free) is free space (cost 0), the plugin itself
    locks the border (reference src/epic_nav_core_plugin.cpp:109-187); goal = the map's goal cell (column 779,
    row 9), start = the first of the golden streamline starts (339, 184)."""
    image = common.maps()["umass"]
    cost = np.where(image == 0, 254, 0).astype(np.uint8)
    path = os.path.join(tmp, "plan_umass_map.bin")
    write_map(path, cost)
    return ["plan", path, None, "779", "9", "%.4f" % (OX + (339 + 0.25) * RES), "%.4f" % (OY + (184 + 0.25) * RES)]


def node_case(tmp):
    """The node on a procedural maze delivered as an OccupancyGrid: 100 = wall, 0 = free, a band of -1
    (unknown, treated as free) and a band of -2 (no change: those cells keep the node's initial state)."""
    u, locked = grids.procedural_maze((120, 160), corridor=8, wall=2, goals=0, seed=7)
    occ = np.where(locked == 1, 100, 0).astype(np.int8)
    occ[40:50, 30:120][occ[40:50, 30:120] == 0] = -1
    occ[100:104, :] = -2
    path = os.path.join(tmp, "node_map.bin")
    write_map(path, occ)
    (gy, gx), (sy, sx) = _far_apart(np.argwhere(occ[1:-1, 1:-1] == 0) + 1)
    return ["node", path, None, "%.4f" % (OX + (gx + 0.5) * RES), "%.4f" % (OY + (gy + 0.5) * RES),
            "%.4f" % (OX + (sx + 0.3) * RES), "%.4f" % (OY + (sy + 0.6) * RES), "60", "50"]


def run(exe, case, process, timeout=300):
    args = [exe] + [process if a is None else a for a in case]
    r = subprocess.run(args, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    out = {}
    for line in r.stdout.splitlines():
        k, _, v = line.partition(" ")
        out[k] = v
    return out
