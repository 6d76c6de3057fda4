"""Shared pieces of the parity tests: the inputs of the golden cases (tests/golden/golden.json was
produced by tests/golden/make_golden.py from the UNTOUCHED reference CPU code) and one checking routine that
every implementation under test goes through -- the oracle, the library's CPU exports, and the CUDA
path through the libepic C ABI.

An implementation is presented as a "solver" with the small interface of oracle.oracle.Oracle:
    .u  .iteration  .delta  run_iterations(n)  complete()  set_cells(v, types)
    potential(x, y)  gradient(x, y, cd)  path(x, y, step, cd, max_length)
"""
import hashlib
import os

import numpy as np

from epic_b200 import grids

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_maps = None


def maps():
    global _maps
    if _maps is None:
        _maps = dict(np.load(os.path.join(ROOT, "tests", "golden", "maps.npz")))
    return _maps


def sha1(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def hexf(v):
    return float(np.float32(v)).hex()


def case_input(name):
    """(u, locked, epsilon, stagger) of a golden case."""
    if name in ("maze", "umass", "basic", "trivial", "c_space"):
        u, locked = grids.grid_from_image(maps()[name])
        return u, locked, 1e-3, 100
    if name == "box64":
        u = np.full((64, 64), -1e6, np.float32)
        locked = np.zeros((64, 64), np.uint32)
        locked[0, :] = locked[-1, :] = locked[:, 0] = locked[:, -1] = 1
        u[10, 10], locked[10, 10] = 0.0, 1
        return u, locked, 1e-3, 100
    if name == "random256":
        return grids.random_obstacles((256, 256), 0.2, 8, 1234) + (1e-3, 100)
    if name == "random_ragged":
        return grids.random_obstacles((131, 77), 0.1, 2, 3) + (1e-3, 100)
    if name == "random48x3":
        return grids.random_obstacles((48, 48, 48), 0.2, 8, 1234) + (1e-3, 100)
    if name == "random3d_ragged":
        return grids.random_obstacles((21, 34, 45), 0.1, 2, 8) + (1e-3, 100)
    if name == "proc_maze":
        return grids.procedural_maze((200, 300), corridor=8, wall=2, goals=3, seed=99) + (1e-3, 100)
    if name == "set_cells":
        return grids.random_obstacles((96, 130), 0.15, 3, 5) + (1e-3, 50)
    raise KeyError(name)


def check_inputs(name, gold):
    u, locked, _, _ = case_input(name)
    if "sha1_u0" in gold:
        assert sha1(u) == gold["sha1_u0"], "generator drifted from the one that made the golden file"
        assert sha1(locked) == gold["sha1_locked"]


def check_checkpoints(make, name, gold, limit=None):
    """Field hash after exactly k iterations of the complete() schedule, for every golden k."""
    u, locked, eps, stagger = case_input(name)
    s = make(u.copy(), locked.copy(), eps, stagger)
    done = 0
    for k in sorted(int(k) for k in gold["checkpoints"]):
        if limit is not None and k > limit:
            break
        s.run_iterations(k - done)
        done = k
        g = gold["checkpoints"][str(k)]
        assert s.iteration == k
        assert sha1(s.u) == g["sha1_u"], "%s: field differs after %d iterations" % (name, k)
        assert hexf(s.delta) == g["delta_hex"], "%s: delta differs after %d iterations" % (name, k)
    return s


def check_complete(make, name, gold):
    """The full solve: iteration count, delta and every bit of the field; then the golden paths."""
    u, locked, eps, stagger = case_input(name)
    s = make(u.copy(), locked.copy(), eps, stagger)
    ret = s.complete()
    g = gold["complete"]
    assert ret == g["ret"]
    assert s.iteration == g["iterations"], "%s: %d iterations, reference %d" % (name, s.iteration, g["iterations"])
    assert hexf(s.delta) == g["delta_hex"]
    assert sha1(s.u) == g["sha1_u"], "%s: converged field is not bit-identical to the reference's" % name
    for (idx, value) in g.get("probes", []):
        assert hexf(s.u[tuple(idx)]) == value
    return s


def check_paths(s, gold):
    for g in gold.get("paths", []):
        ret, p = s.path(g["start"][0], g["start"][1], g["step"], g["cd"], g["max_length"])
        assert ret == g["ret"], "path from %r: return code %d, reference %d" % (g["start"], ret, g["ret"])
        if ret != 0:
            continue
        assert len(p) == g["k"], "path from %r: %d points, reference %d" % (g["start"], len(p), g["k"])
        assert sha1(p) == g["sha1_path"], "path from %r is not bit-identical" % (g["start"],)
        cells = np.floor(p + np.float32(0.5)).astype(np.int64)
        keep = np.ones(len(cells), bool)
        keep[1:] = (cells[1:] != cells[:-1]).any(axis=1)
        assert int(keep.sum()) == g["distinct_cells"]
        assert sha1(cells[keep]) == g["sha1_cells"], "path visits different cells"


def check_potentials(s, gold):
    for g in gold.get("potentials", []):
        x, y = g["xy"]
        ret, v = s.potential(x, y)
        assert ret == g["ret"]
        if ret == 0:
            assert hexf(v) == g["value_hex"]
        ret, gx, gy = s.gradient(x, y, 0.5)
        assert ret == g["grad_ret"]
        if ret == 0:
            assert [hexf(gx), hexf(gy)] == g["grad_hex"]


def check_set_cells(make, gold):
    """Anytime-node style: iterate, edit cells (some out of range / invalid type), iterate on."""
    u, locked, eps, stagger = case_input("set_cells")
    s = make(u.copy(), locked.copy(), eps, stagger)
    s.run_iterations(120)
    ret = s.set_cells(np.array(gold["v"], np.uint32), np.array(gold["types"], np.uint32))
    assert ret == gold["ret"]
    assert sha1(s.u) == gold["sha1_u_after_edit"]
    s.run_iterations(130)
    assert s.iteration == gold["final"]["iterations"]
    assert sha1(s.u) == gold["final"]["sha1_u"]
    assert hexf(s.delta) == gold["final"]["delta_hex"]
    return s


# ---------------------------------------------------------------------------------------------------
# Solvers built on the product library (epic_b200.harmonic.Harmonic -> libepic C ABI)

class LibepicSolver:
    """The libepic C ABI driven the way the reference's callers drive it.  process = 'cpu' exercises the
    library's *_cpu exports, 'gpu' the CUDA path (device-resident between calls, like the anytime node)."""

    def __init__(self, u, locked, epsilon, stagger, process, paths_on="cpu"):
        from epic_b200.harmonic import Harmonic
        self.h = Harmonic(u, locked, epsilon, stagger)
        self.process = process
        self.paths_on = paths_on
        self.resident = False

    def _ensure_resident(self):
        if self.process == "gpu" and not self.resident:
            self.h.initialize_gpu()
            self.resident = True

    @property
    def u(self):
        if self.resident:
            self.h.get_potential_values_gpu()
        return self.h.field

    iteration = property(lambda s: s.h.currentIteration)
    delta = property(lambda s: s.h.delta)

    def run_iterations(self, count):
        self._ensure_resident()
        self.h.run_iterations(count, self.process)

    def complete(self):
        if self.resident:
            self.h.uninitialize_gpu()
            self.resident = False
        self.h.solve(process=self.process)
        return 0

    def set_cells(self, v, types):
        self._ensure_resident()
        r = self.h.set_cells(v, types, "cpu")      # the node edits both copies
        if self.process == "gpu":
            r = self.h.set_cells(v, types, "gpu")
        return r

    def _resident_for_paths(self):
        if self.paths_on == "gpu" and not self.resident:
            self.h.initialize_gpu()
            self.resident = True

    def potential(self, x, y):
        self._resident_for_paths()
        return self.h.compute_potential(x, y, self.paths_on)

    def gradient(self, x, y, cd):
        self._resident_for_paths()
        return self.h.compute_gradient(x, y, cd, self.paths_on)

    def path(self, x, y, step, cd, max_length):
        self._resident_for_paths()
        return self.h.compute_path(x, y, step, cd, max_length, self.paths_on)

    def close(self):
        if self.resident:
            self.h.uninitialize_gpu()
            self.resident = False
