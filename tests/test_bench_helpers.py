"""Host-side pieces of bench.py that carry the correctness evidence of a scaling run: the field hash must not depend
on how the grid is partitioned over ranks, and the generators must hand every rank exactly its rows."""
import numpy as np
import pytest

import bench
from epic_b200 import grids
from epic_b200.sharded import partition


@pytest.mark.parametrize("shape", [(8192, 16), (256, 6, 5)])
def test_field_hash_is_independent_of_the_partition(shape):
    rng = np.random.RandomState(1)
    field = rng.random_sample(shape).astype(np.float32)
    whole = bench.field_hash(bench.band_digests(field, 0))
    for world in (2, 4, 8):
        digests = []
        for rank in range(world):
            row0, rows = partition(shape[0], world, rank)
            assert row0 % bench.hash_band(len(shape)) == 0
            digests += bench.band_digests(field[row0:row0 + rows], row0)
        rng.shuffle(digests)                      # all_gather order must not matter either
        assert bench.field_hash(digests) == whole
    other = field.copy()
    other[shape[0] // 2, 0] += np.float32(1e-3)
    assert bench.field_hash(bench.band_digests(other, 0)) != whole


def test_committed_field_hashes_are_well_formed():
    fields = bench.committed_fields()
    keys = [k for k in fields if not k.startswith("_")]
    assert keys, "tests/golden/bench_fields.json holds the N = 1 hashes the scaling runs are compared with"
    for k in keys:
        assert len(fields[k]["field_sha1"]) == 40 and fields[k]["iterations"] % 100 == 1


@pytest.mark.parametrize("maker,kwargs", [(grids.random_obstacles, dict(p=0.2, goals=5, seed=3)),
                                          (grids.procedural_maze, dict(corridor=6, wall=2, goals=3, seed=3))])
def test_slab_generation_equals_the_whole_grid(maker, kwargs):
    shape = (300, 257)
    u, locked = maker(shape, **kwargs)
    for world in (2, 3, 8):
        for rank in range(world):
            row0, rows = partition(shape[0], world, rank)
            lo, hi = max(0, row0 - 4), min(shape[0], row0 + rows + 4)
            us, ls = maker(shape, row0=lo, rows=hi - lo, **kwargs)
            assert np.array_equal(us, u[lo:hi]) and np.array_equal(ls, locked[lo:hi])


def test_source_stamp_matches_the_committed_traffic_measurement():
    """profiles/ncu_traffic.json is only reported while it describes the kernels in the tree."""
    import json
    import os
    with open(os.path.join(bench.ROOT, "profiles", "ncu_traffic.json")) as f:
        t = json.load(f)
    assert t["source_stamp"] == bench.source_stamp(), "kernel sources changed: re-measure dram bytes with ncu or drop the stamp"
