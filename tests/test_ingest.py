"""map ingest.  The reference's node turns an OccupancyGrid into a set_cells
call over every interior cell and "reset free cells" likewise (src/epic_navigation_node_harmonic.cpp:383-426,
:582-611).  The dense extensions (harmonic_utilities_set_occupancy_grid_2d_*, ..._reset_free_cells_2d_*) must
leave u and locked exactly as that scatter list does -- checked against the oracle's set_cells on the host and
against the host twin on the device, including the sweeps that follow."""
import numpy as np
import pytest

import common
from epic_b200 import grids
from epic_b200.harmonic import Harmonic
from oracle import oracle as orc


def occupancy_message(shape, seed):
    rng = np.random.RandomState(seed)
    occ = rng.randint(0, 101, size=shape).astype(np.int8)
    occ[rng.random_sample(shape) < 0.15] = -1     # unknown: treated as free
    occ[rng.random_sample(shape) < 0.10] = -2     # no change
    return occ


def node_list(u, locked, occ):
    """The (v, types) list the node builds for a /map message."""
    v, types = [], []
    h, w = occ.shape
    for y in range(1, h - 1):
        for x in range(1, w - 1):
            if occ[y, x] == -2 or (u[y, x] == 0.0 and locked[y, x] == 1):
                continue
            v += [x, y]
            types.append(1 if occ[y, x] >= 50 else 2)
    return np.array(v, np.uint32), np.array(types, np.uint32)


def start_state(shape, seed):
    """A node-like state: some goals, some earlier obstacles, relaxed values in the free cells."""
    u, locked = grids.random_obstacles(shape, 0.2, 5, seed=seed)
    o = orc.Oracle(u.copy(), locked.copy(), 1e-3, 100)
    o.run_iterations(37)
    return o.u.copy(), locked


@pytest.mark.parametrize("shape,seed", [((67, 131), 3), ((96, 64), 4)])
def test_host_twins_equal_the_nodes_scatter_lists(libepic_built, shape, seed):
    u, locked = start_state(shape, seed)
    occ = occupancy_message(shape, seed + 100)
    o = orc.Oracle(u.copy(), locked.copy(), 1e-3, 100)
    v, types = node_list(u, locked, occ)
    assert o.set_cells(v, types) == 0
    h = Harmonic(u.copy(), locked.copy(), 1e-3, 100)
    h.set_occupancy_grid(occ, process='cpu')
    assert np.array_equal(h.field, o.u) and np.array_equal(h.locked_cells, o.locked)
    # reset free cells
    free = np.argwhere(o.locked[1:-1, 1:-1] == 0) + 1
    v2 = np.ascontiguousarray(free[:, ::-1], dtype=np.uint32).reshape(-1)
    o.run_iterations(11)
    h2 = Harmonic(o.u.copy(), o.locked.copy(), 1e-3, 100)
    assert o.set_cells(v2, np.full(len(free), 2, np.uint32)) == 0
    h2.reset_free_cells(process='cpu')
    assert np.array_equal(h2.field, o.u) and np.array_equal(h2.locked_cells, o.locked)


def test_ingest_requires_a_resident_field(libepic_built):
    u, locked = start_state((40, 40), 1)
    h = Harmonic(u, locked, 1e-3, 100)
    with pytest.raises(Exception):
        h.set_occupancy_grid(np.zeros((40, 40), np.int8), process='gpu')
    with pytest.raises(Exception):
        h.reset_free_cells(process='gpu')


@pytest.mark.gpu
@pytest.mark.parametrize("shape,seed", [((67, 131), 3), ((300, 700), 4), ((1030, 2051), 5)])
def test_device_ingest_equals_host_twin_and_keeps_sweeping_identically(libepic_built, shape, seed):
    u, locked = start_state(shape, seed)
    occ = occupancy_message(shape, seed + 100)
    h = Harmonic(u.copy(), locked.copy(), 1e-3, 100)
    h.initialize_gpu()
    h.set_occupancy_grid(occ, process='cpu')      # host mirror, as the node keeps one
    h.set_occupancy_grid(occ, process='gpu')
    o = orc.Oracle(h.field.copy(), h.locked_cells.copy(), 1e-3, 100)
    expect_u = h.field.copy()
    h.get_potential_values_gpu()
    assert np.array_equal(h.field, expect_u), "device u after ingest differs from the host twin"
    h.run_iterations(23, "gpu")
    o.run_iterations(23)
    h.get_potential_values_gpu()
    assert np.array_equal(h.field, o.u), "sweeps after ingest differ (free mask on the device is wrong)"
    # reset free cells on both sides, then sweep again
    h.reset_free_cells(process='cpu')
    h.reset_free_cells(process='gpu')
    o2 = orc.Oracle(h.field.copy(), h.locked_cells.copy(), 1e-3, 100)
    o2.iteration = o.iteration
    h.run_iterations(14, "gpu")
    o2.run_iterations(14)
    h.get_potential_values_gpu()
    assert np.array_equal(h.field, o2.u)
    h.uninitialize_gpu()
