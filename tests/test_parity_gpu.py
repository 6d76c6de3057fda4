"""Parity of the CUDA path, through the libepic C ABI, on a B200.

Bit-exactness is the bar in the default (strict) math mode: fields, deltas and iteration counts are
compared with golden vectors produced by the untouched reference CPU code (tests/golden/golden.json)
and with the oracle on seeded inputs; paths are compared point for point.  The fast mode (MUFU
ex2/lg2) is a TOLERANCE mode: same iteration count at matched epsilon, |du| <= 1e-5*|u| + 4e-7*iterations
(test_fast_mode_within_stated_tolerance_at_matched_epsilon explains the second term), streamlines within
0.05 cell of the reference's but not cell-for-cell.
"""
import ctypes as ct
import os

import numpy as np
import pytest

import common
from epic_b200 import grids
from epic_b200 import libepic as le
from epic_b200.field import Field
from epic_b200.harmonic import Harmonic
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

CASES_2D = ["box64", "random256", "random_ragged", "proc_maze", "basic", "umass", "maze"]
CASES_3D = ["random48x3", "random3d_ragged"]


def make_gpu(u, locked, eps, stagger):
    return common.LibepicSolver(u, locked, eps, stagger, "gpu", paths_on="gpu")


def make_gpu_host_paths(u, locked, eps, stagger):
    return common.LibepicSolver(u, locked, eps, stagger, "gpu", paths_on="cpu")


def test_strict_math_device_functions_equal_host_libm(libepic_built):
    """Every float the sweep can feed to expf (x <= 0) and logf ([1, 8]): device result == host glibc."""
    vals = [ct.c_uint64(0) for _ in range(4)]
    assert libepic_built.epic_b200_selftest_math(1, *[ct.byref(v) for v in vals]) == 0
    assert vals[0].value == 0xff800000 - 0x80000000 + 1 and vals[2].value == 0x41000000 - 0x3f800000 + 1
    assert vals[1].value == 0, "%d expf mismatches" % vals[1].value
    assert vals[3].value == 0, "%d logf mismatches" % vals[3].value


@pytest.mark.parametrize("name", CASES_2D + CASES_3D)
def test_checkpoints_bit_exact(golden, libepic_built, name):
    """update / update_and_check call by call (the anytime node's usage): field and delta after k iterations."""
    s = common.check_checkpoints(make_gpu, name, golden[name])
    s.close()


@pytest.mark.parametrize("name", CASES_2D + CASES_3D)
def test_complete_gpu_bit_exact_and_paths(golden, libepic_built, name):
    """harmonic_complete_gpu to epsilon: iterations, delta, every bit of the field; then the golden
    paths extracted on the device-resident field (cell-for-cell and point-for-point)."""
    s = common.check_complete(make_gpu, name, golden[name])
    if s.h.n == 2:
        common.check_paths(s, golden[name])
        common.check_potentials(s, golden[name])
    s.close()


def test_plugin_sequence_host_paths(golden, libepic_built):
    """nav_core plugin order (src/epic_nav_core_plugin.cpp:256-298): complete_gpu, then
    harmonic_compute_path_2d_cpu on the host copy of u."""
    s = common.check_complete(make_gpu_host_paths, "umass", golden["umass"])
    common.check_paths(s, golden["umass"])


def test_set_cells_sequence(golden, libepic_built):
    s = common.check_set_cells(make_gpu, golden["set_cells"])
    s.close()


def test_set_cells_duplicates_last_edit_wins(libepic_built):
    u, locked, eps, stagger = common.case_input("box64")
    a = common.LibepicSolver(u.copy(), locked.copy(), eps, stagger, "gpu")
    o = orc.Oracle(u.copy(), locked.copy(), eps, stagger)
    v = np.array([[5, 5], [5, 5], [6, 6], [6, 6], [5, 5], [70, 3], [3, 70]], np.uint32)
    t = np.array([0, 1, 1, 2, 2, 0, 0], np.uint32)
    a.run_iterations(3)
    o.run_iterations(3)
    assert a.set_cells(v, t) == 0 and o.set_cells(v, t) == 0
    a.run_iterations(50)
    o.run_iterations(50)
    assert np.array_equal(a.u, o.u)
    a.close()


def test_python_wrapper_double_initialise_sequence(golden, libepic_built):
    """The reference's Harmonic.solve (python/epic/harmonic.py:73-97) initialises, calls complete_gpu
    (which initialises again) and uninitialises: must work and leave null handles."""
    u, locked, eps, stagger = common.case_input("box64")
    h = Harmonic(u, locked, eps, stagger)
    L = libepic_built
    assert L.harmonic_initialize_dimension_size_gpu(ct.byref(h)) == 0
    assert L.harmonic_initialize_potential_values_gpu(ct.byref(h)) == 0
    assert L.harmonic_initialize_locked_gpu(ct.byref(h)) == 0
    assert L.harmonic_complete_gpu(ct.byref(h), 1024) == 0
    assert not h.d_m and not h.d_u and not h.d_locked and not h.d_delta
    assert L.harmonic_uninitialize_dimension_size_gpu(ct.byref(h)) == 0
    assert L.harmonic_uninitialize_potential_values_gpu(ct.byref(h)) == 0
    assert L.harmonic_uninitialize_locked_gpu(ct.byref(h)) == 0
    assert h.currentIteration == golden["box64"]["complete"]["iterations"]
    assert common.sha1(h.field) == golden["box64"]["complete"]["sha1_u"]


def test_execute_gpu_error_codes(libepic_built):
    u, locked, eps, stagger = common.case_input("box64")
    h = Harmonic(u, locked, eps, stagger)
    h.initialize_gpu()
    L = libepic_built
    assert L.harmonic_execute_gpu(ct.byref(h), 100) == le.EPIC_ERROR_INVALID_CUDA_PARAM
    h.epsilon = 0.0
    assert L.harmonic_execute_gpu(ct.byref(h), 1024) == le.EPIC_ERROR_INVALID_DATA
    h.epsilon = 1e-3
    assert L.harmonic_initialize_gpu(ct.byref(h), 1024) == 0
    assert L.harmonic_execute_gpu(ct.byref(h), 1024) == le.EPIC_ERROR_INVALID_DATA   # d_delta already set
    assert L.harmonic_uninitialize_gpu(ct.byref(h)) == 0
    assert L.harmonic_execute_gpu(ct.byref(h), 1024) == 0
    h.uninitialize_gpu()


@pytest.mark.parametrize("k", [1, 2, 3, 5])
def test_update_gpu_then_execute_gpu_keeps_the_queued_sweeps(libepic_built, k):
    """harmonic_update_gpu x k (k not a multiple of the pass depth T = 4, so sweeps are still queued inside the
    library) followed by harmonic_execute_gpu: the reference has applied those sweeps to d_u before execute
    resets currentIteration (harmonic_gpu.cu:226-262), so the solve must start from the swept field."""
    u, locked, eps, stagger = common.case_input("random_ragged")
    h = Harmonic(u.copy(), locked.copy(), eps, stagger)
    h.initialize_gpu()
    L = libepic_built
    for _ in range(k):
        assert L.harmonic_update_gpu(ct.byref(h), 1024) == 0
    assert h.currentIteration == k
    o = orc.Oracle(u.copy(), locked.copy(), eps, stagger)
    o.run_iterations(k)          # iterations 0 .. k-1: the same cells, whether or not delta is taken
    o.iteration = 0
    assert o.complete() == 0
    assert L.harmonic_execute_gpu(ct.byref(h), 1024) == 0
    assert h.currentIteration == o.iteration
    assert h.delta == o.delta
    assert np.array_equal(h.field, o.u)
    h.uninitialize_gpu()


def test_update_model_reuploads(libepic_built):
    u, locked, eps, stagger = common.case_input("random_ragged")
    s = common.LibepicSolver(u.copy(), locked.copy(), eps, stagger, "gpu")
    o = orc.Oracle(u.copy(), locked.copy(), eps, stagger)
    s.run_iterations(7)
    # host-side edit + harmonic_update_model_gpu (reference harmonic_model_gpu.cu:172-204)
    s.h.field[:] = u
    s.h.locked_cells[40:60, 20:30] = 1
    o.locked[40:60, 20:30] = 1
    s.h.currentIteration = 0
    s.h.update_model_gpu()
    s.run_iterations(33)
    o.run_iterations(33)
    assert np.array_equal(s.u, o.u)
    s.close()


@pytest.mark.parametrize("shape,p,goals", [((700, 1100), 0.2, 5), ((1030, 517), 0.0, 1), ((300, 4099), 0.35, 9),
                                           ((2048, 2048), 0.2, 16)])
def test_seeded_grids_against_oracle(libepic_built, shape, p, goals):
    """Shapes that straddle tile and pitch boundaries; fixed-K comparison with the oracle."""
    u, locked = grids.random_obstacles(shape, p, goals, seed=77)
    s = common.LibepicSolver(u.copy(), locked.copy(), 1e-3, 100, "gpu")
    o = orc.Oracle(u.copy(), locked.copy(), 1e-3, 100, threads=8)
    for k in (1, 2, 98, 1, 27):
        s.run_iterations(k)
        o.run_iterations(k)
        assert s.iteration == o.iteration
        assert np.array_equal(s.u, o.u), "field differs after %d iterations" % o.iteration
        assert s.delta == o.delta
    s.close()


def test_seeded_3d_against_oracle(libepic_built):
    u, locked = grids.random_obstacles((70, 45, 150), 0.2, 6, seed=5)
    s = common.LibepicSolver(u.copy(), locked.copy(), 1e-3, 100, "gpu")
    o = orc.Oracle(u.copy(), locked.copy(), 1e-3, 100, threads=8)
    for k in (1, 2, 98, 1, 9):
        s.run_iterations(k)
        o.run_iterations(k)
        assert np.array_equal(s.u, o.u) and s.delta == o.delta
    s.close()


def test_unlocked_border_is_never_updated(libepic_built):
    """The reference sweeps interior cells only (harmonic_cpu.cpp:46-51), whatever `locked` says."""
    u, locked = grids.random_obstacles((90, 300), 0.1, 3, seed=2)
    locked[0, :] = locked[-1, :] = locked[:, 0] = locked[:, -1] = 0
    s = common.LibepicSolver(u.copy(), locked.copy(), 1e-3, 100, "gpu")
    o = orc.Oracle(u.copy(), locked.copy(), 1e-3, 100)
    s.run_iterations(40)
    o.run_iterations(40)
    assert np.array_equal(s.u, o.u)
    s.close()


FAST_TOL_REL, FAST_TOL_PER_ITERATION = 1e-5, 4e-7


@pytest.mark.parametrize("case", ["random256", "proc_maze", "basic", "umass", "maze", "random48x3"])
def test_fast_mode_within_stated_tolerance_at_matched_epsilon(libepic_built, golden, case):
    """EPIC_MATH=fast (MUFU ex2/lg2, the arithmetic of the reference's own GPU kernel) solved to the same
    epsilon as the reference CPU path: SAME iteration count, and converged log-potentials within

        |u_fast - u_ref| <= 1e-5 * |u_ref| + 4e-7 * iterations.

    The second term is the price of lg2.approx: its absolute error (up to 2^-22 on [1, 4]) enters every
    update directly, and the slowly converging modes of the relaxation do not damp it, so the offset grows
    with the iteration count (measured: 8.6e-3 on umass.png after 32 701 iterations, 2.3e-3 on maze.png
    after 49 301).  It is a smooth offset: streamlines move by < 0.03 cell (next test).  u_ref is the strict
    GPU field, which the other tests pin bit for bit to the reference (its sha1 is re-checked here)."""
    u, locked, eps, stagger = common.case_input(case)
    res = {}
    for math in ("strict", "fast"):
        f = Field(u.shape, math=math)
        f.upload(u, locked)
        it, d = f.solve(eps, stagger)
        res[math] = (it, d, f.download_u())
        f.close()
    g = golden[case]["complete"]
    assert res["strict"][0] == g["iterations"] and common.sha1(res["strict"][2]) == g["sha1_u"]
    assert res["fast"][0] == res["strict"][0], "iteration counts at matched epsilon differ"
    free = locked == 0
    ref, got = res["strict"][2], res["fast"][2]
    err = np.abs(got[free] - ref[free])
    tol = FAST_TOL_REL * np.abs(ref[free]) + FAST_TOL_PER_ITERATION * res["fast"][0]
    assert np.all(err <= tol), "%d cells outside the tolerance, max err %g" % (int((err > tol).sum()), err.max())
    assert np.array_equal(got[~free], ref[~free]), "locked cells must be untouched"
    assert abs(res["fast"][1] - res["strict"][1]) <= 1e-5


@pytest.mark.parametrize("case", ["umass", "maze", "basic"])
def test_fast_mode_streamlines_stay_within_a_fraction_of_a_cell(libepic_built, golden, case):
    """Fast-mode streamlines against the reference's (strict field, bit-identical to the CPU path): same
    return code, length within 2 points, every point within 0.05 cell.  (Cell-for-cell identity is a
    strict-mode guarantee only: streamlines run along corridor axes that lie on cell boundaries, where a
    1e-6 perturbation flips the nearest cell.)"""
    u, locked, eps, stagger = common.case_input(case)
    starts = [p["start"] for p in golden[case]["paths"] if p["step"] == 0.05]
    starts += [list(map(float, s)) for s in grids.free_cells(locked, 60, seed=11)]
    paths = {}
    for math in ("strict", "fast"):
        f = Field(u.shape, math=math)
        f.upload(u, locked)
        f.solve(eps, stagger)
        paths[math] = f.paths(starts, 0.05, 0.5, int(u.size / 0.05))
        f.close()
    for (ra, pa), (rb, pb) in zip(paths["strict"], paths["fast"]):
        assert ra == rb
        if ra == 0:
            assert abs(len(pa) - len(pb)) <= 2
            n = min(len(pa), len(pb))
            assert np.abs(pa[:n] - pb[:n]).max() <= 0.05


def test_fast_mode_fixed_iterations_against_oracle(libepic_built):
    """Fast mode against the CPU oracle itself at a fixed iteration count (shallow field, few iterations)."""
    u, locked, eps, stagger = common.case_input("random256")
    f = Field(u.shape, math="fast")
    f.upload(u, locked)
    o = orc.Oracle(u.copy(), locked.copy(), eps, stagger)
    f.run(0, 601, check_last=True)
    d = f.read_delta()
    o.run_iterations(601)
    got = f.download_u()
    free = locked == 0
    err = np.abs(got[free] - o.u[free])
    assert np.all(err <= FAST_TOL_REL * np.abs(o.u[free]) + FAST_TOL_PER_ITERATION * 601), "max err %g" % err.max()
    assert abs(d - o.delta) <= 1e-5
    f.close()


def test_field_solve_matches_execute(libepic_built, golden):
    u, locked, eps, stagger = common.case_input("proc_maze")
    f = Field(u.shape)
    f.upload(u, locked)
    it, d = f.solve(eps, stagger)
    g = golden["proc_maze"]["complete"]
    assert it == g["iterations"] and common.hexf(d) == g["delta_hex"]
    assert common.sha1(f.download_u()) == g["sha1_u"]
    assert np.array_equal(f.download_locked(), locked)
    f.close()


def test_batched_paths_equal_single_paths(libepic_built, golden):
    s = common.check_complete(make_gpu, "basic", golden["basic"])
    s._resident_for_paths()
    starts = grids.free_cells(s.h.locked_cells, 40, seed=3)
    batch = s.h.compute_paths_gpu(starts, 0.05, 0.5, int(s.h.field.size / 0.05))
    s.h.get_potential_values_gpu()
    for (x, y), (ret, p) in zip(starts, batch):
        r2, p2 = s.h.compute_path(float(x), float(y), 0.05, 0.5, int(s.h.field.size / 0.05), "cpu")
        assert ret == r2 and np.array_equal(p, p2)
    s.close()


def test_round_trip_and_idempotence_at_full_size(libepic_built):
    """16384^2 (BASELINE.json config 3) through size-independent properties: upload/download round trip,
    locked cells never change, a sweep of one colour leaves the other colour untouched, and the delta of a
    check sweep equals the max change observed from outside."""
    shape = (16384, 16384)
    u, locked = grids.random_obstacles(shape, 0.2, 64, seed=1234)
    f = Field(shape)
    f.upload(u, locked)
    assert np.array_equal(f.download_u(first=8000, layers=64), u[8000:8064])
    f.run(0, 1, check_last=True)
    d = f.read_delta()
    a = f.download_u()
    changed = a != u
    yy, xx = np.nonzero(changed)
    assert np.all((yy + xx) % 2 == 1), "iteration 0 may only touch cells with (row + column) odd"
    assert not changed[locked != 0].any()
    assert d == np.abs(a[changed] - u[changed]).max()
    f.run(1, 12, check_last=False)
    b = f.download_u()
    assert np.array_equal(b[locked != 0], u[locked != 0])
    # rows 0..47 against the oracle on a band (the band's last rows lack their lower neighbours, so
    # only rows that 13 sweeps cannot reach from the cut are compared)
    band = 48 + 13
    o = orc.Oracle(u[:band].copy(), np.where(np.arange(band)[:, None] == band - 1, 1, locked[:band]).astype(np.uint32),
                   1e-3, 100, threads=8)
    o.run_iterations(13)
    assert np.array_equal(b[:48], o.u[:48])
    f.close()


def _random_state_band(shape, lo, hi, seed):
    """Rows / layers [lo, hi) of a grid whose free cells hold arbitrary relaxed-looking values, so that every
    cell moves from the first sweep on (a field of -1e6 with a few goals would leave most of a band untouched)."""
    u, locked = grids.random_obstacles(shape, 0.2, 0, seed=seed, row0=lo, rows=hi - lo)
    rng = np.random.RandomState([seed, lo])
    vals = -40.0 * rng.random_sample(u.shape).astype(np.float32)
    u = np.where(locked == 0, vals, u).astype(np.float32)
    goals = (rng.random_sample(u.shape) < 1e-4) & (locked == 0)
    inner = np.zeros(u.shape, bool)
    inner[(slice(1, -1),) * u.ndim] = True
    goals &= inner
    u[goals] = 0.0
    locked[goals] = 1
    return u, locked


def _check_bands(shape, bands, sweeps, seed):
    f = Field(shape)
    held = {}
    for lo, hi in bands:
        held[lo] = _random_state_band(shape, lo, hi, seed)
        f.upload(*held[lo], first=lo, layers=hi - lo)
    f.run(0, sweeps, check_last=False)
    for lo, hi in bands:
        u, locked = held[lo]
        got = f.download_u(first=lo, layers=hi - lo)
        # the oracle sees the band as a grid of its own: its first / last layers are never updated there,
        # while on the device they are interior cells fed by the (empty) rest of the field -- compare the
        # layers that `sweeps` sweeps cannot reach from such a cut; a cut on the global border is exact
        o = orc.Oracle(u.copy(), locked.copy(), 1e-3, 1000, threads=8)
        assert lo % 2 == 0, "the colour phase follows the global layer index"
        o.run_iterations(sweeps)
        a = 0 if lo == 0 else sweeps
        b = (hi - lo) if hi == shape[0] else (hi - lo) - sweeps
        assert np.array_equal(got[a:b], o.u[a:b]), "band %d..%d of %s differs from the oracle" % (lo, hi, shape)
        assert not np.array_equal(got[a:b], u[a:b])
    f.close()


def test_maximum_size_2d_sampled_bands_against_oracle(libepic_built):
    """65536 x 65536 (BASELINE.json config 4's size): N = 2^32 cells, one more than the reference's unsigned
    int can count, 33 GiB resident.  Three row bands (top border, middle, bottom border) hold a
    seeded state, the rest of the field stays empty; 12 half-sweeps (three passes) on the device must equal the
    64-bit oracle on each band -- 64-bit addressing, tile indexing over 197 000 CTAs, border handling at row
    65535."""
    size = 65536
    _check_bands((size, size), [(0, 192), (size // 2 - 96, size // 2 + 96), (size - 192, size)], 12, seed=77)


def test_maximum_size_3d_sampled_slabs_against_oracle(libepic_built):
    """1024^3 (BASELINE.json config 5): x0-slabs at the top border, in the middle and at the bottom border."""
    _check_bands((1024, 1024, 1024), [(0, 20), (510, 530), (1004, 1024)], 6, seed=78)


@pytest.mark.parametrize("shape", [(1536, 2048), (200, 230, 260)])
def test_static_tile_skipping_is_bit_identical_and_active(libepic_built, monkeypatch, shape):
    """Field::solve skips tiles (3-D: column chunks) whose neighbourhood saw no update change a value in the
    previous pass (the replayed computation is a no-op into a buffer that already holds the data).  Same
    iteration count, delta and field with the feature off, and the feature must actually engage on a grid whose
    wave fronts take a while to fill it."""
    u, locked = grids.random_obstacles(shape, 0.2, 3, seed=9)
    out = {}
    for skip in ("0", "1"):
        monkeypatch.setenv("EPIC_SKIP_STATIC", skip)
        f = Field(shape)
        f.upload(u, locked)
        it, delta = f.solve(1e-3, 100)
        out[skip] = (it, delta, common.sha1(f.download_u()), f.info()["skipped_tiles"])
        # a second solve on the converged field: nearly everything is static from the second pass on
        it2, delta2 = f.solve(1e-3, 100)
        out[skip] += (it2, delta2, common.sha1(f.download_u()), f.info()["skipped_tiles"])
        f.close()
    assert out["0"][:3] == out["1"][:3] and out["0"][4:7] == out["1"][4:7]
    assert out["0"][3] == 0 and out["0"][7] == 0
    assert out["1"][3] > 0 and out["1"][7] > out["1"][3]


@pytest.mark.parametrize("sweeps_per_pass,tile_rows,threads", [
    (1, 16, 256), (2, 24, 256), (3, 32, 512), (4, 16, 256), (4, 40, 256), (4, 96, 512), (4, 64, 512), (5, 48, 256),
    (6, 56, 512), (8, 32, 256), (8, 96, 512)])
def test_every_tile_geometry_gives_the_same_bits(golden, libepic_built, monkeypatch, sweeps_per_pass, tile_rows, threads):
    """The cost model picks (threads, tile rows) from the grid size and T defaults to 4; every other legal
    geometry -- halo depth 1..8, 16..96 tile rows, 8 or 16 warps -- must produce the golden fields too:
    a ragged random grid call by call, and a procedural maze solved to epsilon (with static-tile skipping)."""
    monkeypatch.setenv("EPIC_SWEEPS_PER_PASS", str(sweeps_per_pass))
    monkeypatch.setenv("EPIC_TILE_ROWS", str(tile_rows))
    monkeypatch.setenv("EPIC_THREADS", str(threads))
    s = common.check_checkpoints(make_gpu, "random_ragged", golden["random_ragged"])
    s.close()
    s = common.check_complete(make_gpu, "proc_maze", golden["proc_maze"])
    s.close()


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (3, 3), (1, 50), (50, 1), (2, 7), (3, 300), (300, 3), (4, 4),
                                   (1, 1, 1), (3, 3, 3), (2, 5, 9), (5, 2, 9), (9, 5, 2), (4, 4, 4)])
def test_degenerate_grids_match_the_oracle(libepic_built, shape):
    """Grids with no interior (a dimension below 3) are legal inputs of the reference: its loops simply do not
    run.  The device path must agree: nothing changes, delta is 0, the iteration counter still advances; with
    the smallest interiors (3 and 4 cells across) the few free cells follow the oracle bit for bit."""
    rng = np.random.RandomState(sum(shape) * 7 + len(shape))
    u = (-30.0 * rng.random_sample(shape)).astype(np.float32)
    locked = (rng.random_sample(shape) < 0.3).astype(np.uint32)
    u[locked == 1] = np.where(rng.random_sample(int(locked.sum())) < 0.5, 0.0, -1e6).astype(np.float32)
    h = Harmonic(u.copy(), locked.copy(), 1e-3, 3)
    o = orc.Oracle(u.copy(), locked.copy(), 1e-3, 3)
    h.initialize_gpu()
    h.run_iterations(7, "gpu")
    o.run_iterations(7)
    h.get_potential_values_gpu()
    h.uninitialize_gpu()
    assert h.currentIteration == o.iteration == 7
    assert np.array_equal(h.field, o.u)
    assert h.delta == o.delta
