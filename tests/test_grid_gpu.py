"""Multi-GPU behind the libepic C ABI (epic_b200/csrc/engine/grid.cu): with EPIC_DEVICES set, the reference's
own entry points -- harmonic_complete_gpu, the update / update_and_check pair, set_cells, the streamline calls --
shard the grid over several slabs inside the library.  Red-black ordering makes the result independent of the
partition, so every golden vector (made by the untouched reference CPU code) must be reproduced bit for bit.

On a one-GPU box the slabs share the device (EPIC_DEVICES=0,0,...): everything slab-specific runs -- ghost layers,
peer stores and in-kernel pass ordering, the all-reduce of the convergence check inside the decision kernels, the
multi-slab streamline view -- only NVLink is missing.  With >= 2 GPUs the same tests also run on distinct devices
(one host thread per slab)."""
import ctypes as ct

import numpy as np
import pytest
import torch

import common
import replay
from epic_b200 import grids
from epic_b200 import libepic as le
from epic_b200.harmonic import Harmonic
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def device_lists():
    lists = ["0,0", "0,0,0"]
    n = torch.cuda.device_count()
    if n >= 2:
        lists.append(",".join(str(i) for i in range(min(n, 8))))
        lists.append("0,1")
    return lists


@pytest.mark.parametrize("devices", device_lists())
@pytest.mark.parametrize("case", ["box64", "basic", "proc_maze", "random256", "random48x3"])
def test_complete_gpu_sharded_behind_the_abi(golden, libepic_built, monkeypatch, devices, case):
    monkeypatch.setenv("EPIC_DEVICES", devices)
    monkeypatch.setenv("EPIC_MIN_SLAB_CELLS", "0")     # take the list literally, however small the grid
    u, locked, eps, stagger = common.case_input(case)
    h = Harmonic(u.copy(), locked.copy(), eps, stagger)
    assert libepic_built.harmonic_complete_gpu(ct.byref(h), 1024) == 0
    g = golden[case]["complete"]
    assert h.currentIteration == g["iterations"]
    assert common.hexf(h.delta) == g["delta_hex"]
    assert common.sha1(h.field) == g["sha1_u"]


@pytest.mark.parametrize("devices", device_lists()[:1] + device_lists()[2:3])
def test_checkpoints_and_paths_sharded_behind_the_abi(golden, libepic_built, monkeypatch, devices):
    """update / update_and_check call by call (host-side max of the slabs' deltas), then streamlines on the
    device-resident sharded field: the kernels read the slabs through a multi-slab view."""
    monkeypatch.setenv("EPIC_DEVICES", devices)
    monkeypatch.setenv("EPIC_MIN_SLAB_CELLS", "0")     # take the list literally, however small the grid
    make = lambda u, l, e, s: common.LibepicSolver(u, l, e, s, "gpu")   # noqa: E731
    common.check_checkpoints(make, "proc_maze", golden["proc_maze"])
    u, locked, eps, stagger = common.case_input("box64")
    s = common.LibepicSolver(u.copy(), locked.copy(), eps, stagger, "gpu", paths_on="gpu")
    o = orc.Oracle(u.copy(), locked.copy(), eps, stagger)
    s.complete()
    o.complete()
    assert s.iteration == o.iteration
    for start in ((50.0, 50.0), (12.0, 55.5), (40.25, 20.0)):
        rp, p = s.path(*start, 0.05, 0.5, 81920)
        ro, po = o.path(*start, 0.05, 0.5, 81920)
        assert rp == ro and np.array_equal(p, po)
        rv, v = s.potential(*start)
        rov, ov = o.potential(*start)
        assert (rv, v) == (rov, ov)
    s.close()


@pytest.mark.parametrize("devices", device_lists()[:1] + device_lists()[2:3])
def test_set_cells_and_update_model_sharded(libepic_built, monkeypatch, devices):
    """Sparse edits (also into ghost layers: rows next to a slab boundary) and a full re-upload."""
    monkeypatch.setenv("EPIC_DEVICES", devices)
    monkeypatch.setenv("EPIC_MIN_SLAB_CELLS", "0")     # take the list literally, however small the grid
    shape = (160, 130)
    u, locked = grids.random_obstacles(shape, 0.15, 3, seed=5)
    s = common.LibepicSolver(u.copy(), locked.copy(), 1e-3, 50, "gpu")
    o = orc.Oracle(u.copy(), locked.copy(), 1e-3, 50)
    rng = np.random.RandomState(3)
    for round_ in range(4):
        s.run_iterations(37)
        o.run_iterations(37)
        k = 60
        v = np.stack([rng.randint(0, shape[1] + 2, k), rng.randint(70, 92, k)], 1).astype(np.uint32)   # around row 80
        types = rng.randint(0, 3, k).astype(np.uint32)
        interior = (v[:, 0] > 0) & (v[:, 0] < shape[1] - 1)
        v, types = v[interior], types[interior]
        s.set_cells(v, types)
        o.set_cells(v, types)
    s.run_iterations(50)
    o.run_iterations(50)
    assert np.array_equal(s.u, o.u) and s.delta == o.delta
    s.close()


@pytest.mark.parametrize("devices", device_lists()[:1] + device_lists()[2:3])
@pytest.mark.parametrize("scenario", ["plan", "node"])
def test_replayed_callers_with_epic_devices(libepic_built, monkeypatch, tmp_path, devices, scenario):
    """The ROS-free replay of the plugin / the node, unchanged, with the library sharding behind the ABI."""
    import json
    monkeypatch.setenv("EPIC_DEVICES", devices)
    monkeypatch.setenv("EPIC_MIN_SLAB_CELLS", "0")     # take the list literally, however small the grid
    exe = replay.build(str(tmp_path / "replay_ours"))
    with open(replay.GOLDEN) as f:
        gold = json.load(f)[scenario]
    case = replay.plan_case(str(tmp_path)) if scenario == "plan" else replay.node_case(str(tmp_path))
    out = replay.run(exe, case, "gpu")
    for k, v in gold.items():
        assert out.get(k) == v, "%s: got %s, the reference gives %s" % (k, out.get(k), v)


def test_solve_after_solve_and_updates_after_solve(libepic_built, monkeypatch):
    """Passes queued past the converged check retire as no-ops but still publish their index to the neighbours:
    the next calls on the same resident grid must neither hang nor differ."""
    monkeypatch.setenv("EPIC_DEVICES", "0,0")
    monkeypatch.setenv("EPIC_MIN_SLAB_CELLS", "0")
    u, locked, eps, stagger = common.case_input("random_ragged")
    h = Harmonic(u.copy(), locked.copy(), eps, stagger)
    h.initialize_gpu()
    L = libepic_built
    o = orc.Oracle(u.copy(), locked.copy(), eps, stagger)
    assert o.complete() == 0
    assert L.harmonic_execute_gpu(ct.byref(h), 1024) == 0
    assert h.currentIteration == o.iteration and np.array_equal(h.field, o.u)
    h.run_iterations(9, "gpu")          # continues from currentIteration
    o.run_iterations(9)
    h.get_potential_values_gpu()
    assert np.array_equal(h.field, o.u)
    o.iteration = 0
    assert o.complete() == 0
    assert L.harmonic_execute_gpu(ct.byref(h), 1024) == 0
    assert h.currentIteration == o.iteration and np.array_equal(h.field, o.u)
    h.uninitialize_gpu()
