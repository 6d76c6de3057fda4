"""Pins the oracle (oracle/harmonic_oracle.c, the CPU restatement used as the checker everywhere else)
(a) against tests/golden/golden.json, produced by the untouched reference CPU code, and
(b) against that reference itself where oracle/_ref/libepic_ref_cpu.so exists (the build container;
    /root/reference is not present on the GPU box)."""
import numpy as np
import pytest

import common
from oracle import oracle as orc

FAST_FULL = ["box64", "random256", "random_ragged", "random48x3", "random3d_ragged", "proc_maze"]
ALL = FAST_FULL + ["basic", "umass", "maze"]


def make_oracle(u, locked, eps, stagger):
    return orc.Oracle(u, locked, eps, stagger)


@pytest.mark.parametrize("name", ALL)
def test_oracle_checkpoints_match_reference_golden(golden, name):
    common.check_inputs(name, golden[name])
    common.check_checkpoints(make_oracle, name, golden[name])


@pytest.mark.parametrize("name", FAST_FULL)
def test_oracle_full_solve_matches_reference_golden(golden, name):
    s = common.check_complete(make_oracle, name, golden[name])
    common.check_paths(s, golden[name])
    common.check_potentials(s, golden[name])


def test_oracle_set_cells_matches_reference_golden(golden):
    s = common.check_set_cells(make_oracle, golden["set_cells"])
    assert common.sha1(s.locked) == golden["set_cells"]["sha1_locked_final"]


def test_oracle_threads_do_not_change_bits(golden):
    u, locked, eps, stagger = common.case_input("random256")
    a = orc.Oracle(u.copy(), locked.copy(), eps, stagger, threads=1)
    b = orc.Oracle(u.copy(), locked.copy(), eps, stagger, threads=4)
    a.run_iterations(101)
    b.run_iterations(101)
    assert np.array_equal(a.u, b.u) and a.delta == b.delta


def test_oracle_rejects_bad_input():
    u, locked, _, _ = common.case_input("box64")
    assert orc.Oracle(u.copy(), locked.copy(), 0.0, 100).complete() == orc.INVALID_DATA
    assert orc.Oracle(u.copy(), locked.copy(), -1.0, 100).complete() == orc.INVALID_DATA
    o = orc.Oracle(u.copy(), locked.copy(), 1e-3, 100)
    assert o.set_cells(np.zeros((0, 2), np.uint32), np.zeros(0, np.uint32)) == orc.INVALID_DATA


@pytest.mark.skipif(not orc.have_ref(), reason="needs /root/reference (build container only)")
@pytest.mark.parametrize("name", ["box64", "random_ragged", "random3d_ragged", "c_space", "trivial"])
def test_oracle_equals_reference_build(name):
    u, locked, eps, stagger = common.case_input(name)
    a = orc.Oracle(u.copy(), locked.copy(), eps, stagger)
    b = orc.Reference(u.copy(), locked.copy(), eps, stagger)
    for k in (1, 1, 5, 94, 1, 60):
        a.run_iterations(k)
        b.run_iterations(k)
        assert a.iteration == b.iteration and np.array_equal(a.u, b.u)
        assert a.delta == b.delta
    if u.ndim == 2:
        rng = np.random.RandomState(5)
        for _ in range(200):
            x, y = rng.uniform(-1, u.shape[1] + 1), rng.uniform(-1, u.shape[0] + 1)
            # the reference indexes outside its arrays for points within half a cell of the edge
            if not (1.0 <= x <= u.shape[1] - 2.0 and 1.0 <= y <= u.shape[0] - 2.0):
                continue
            assert a.potential(x, y) == b.potential(x, y)
            ga, gb = a.gradient(x, y, 0.5), b.gradient(x, y, 0.5)
            assert ga[0] == gb[0]
            if ga[0] == 0:
                assert np.array_equal(np.float32(ga[1:]), np.float32(gb[1:]), equal_nan=True)
