"""torchrun worker of tests/test_sharded_gpu.py::test_maximum_size_2d_sharded_*: BASELINE.json config 4's grid size,
65536 x 65536 = 2^32 cells (one more than the reference's `unsigned int` can count), row-sharded
over all ranks with NVLink peer stores for the halos.  Row bands at the global borders and STRADDLING every slab
boundary hold a seeded relaxed-looking state (the rest of the field stays locked and empty, so nothing but the bands
has to cross PCIe); after 12 half-sweeps (three passes, i.e. three halo exchanges) every band must equal the 64-bit
oracle run on that band alone -- 64-bit addressing inside slabs of up to 2^31 cells, ghost-row exchange at those
offsets, the colour phase of a slab that starts at row 32768, border handling at row 65535."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from epic_b200.sharded import GpuSlab, ShardedSolver, partition  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from test_parity_gpu import _random_state_band  # noqa: E402

size, out = int(sys.argv[1]), sys.argv[2]
sweeps, half = 12, 96
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
shape = (size, size)
bands = [(0, 2 * half)] + [(partition(size, world, r)[0] - half, partition(size, world, r)[0] + half) for r in range(1, world)] + \
        [(size - 2 * half, size)]
slab = GpuSlab(shape, rank, world, halo="p2p")
lo, hi = slab.held_range()
for a, b in bands:
    x, y = max(a, lo), min(b, hi)
    if x < y:
        u, locked = _random_state_band(shape, a, b, 77)
        slab.field.upload(u[x - a:y - a], locked[x - a:y - a], first=x, layers=y - x)
torch.cuda.synchronize()
dist.barrier()
solver = ShardedSolver(slab)
solver.run(sweeps, False)
torch.cuda.synchronize()
checked, bad = 0, []
own_lo, own_hi = slab.row0, slab.row0 + slab.rows
for a, b in bands:
    x, y = max(a, own_lo), min(b, own_hi)
    if x >= y:
        continue
    u, locked = _random_state_band(shape, a, b, 77)
    o = orc.Oracle(u.copy(), locked.copy(), 1e-3, 1000, threads=8)
    assert a % 2 == 0
    o.run_iterations(sweeps)
    # rows the cut of the band cannot have reached in `sweeps` sweeps (a cut on the global border is exact)
    va = a if a == 0 else a + sweeps
    vb = b if b == size else b - sweeps
    x, y = max(x, va), min(y, vb)
    got = slab.field.download_u(first=x, layers=y - x)
    checked += y - x
    if not np.array_equal(got, o.u[x - a:y - a]):
        bad.append([a, b, x, y])
    if np.array_equal(got, u[x - a:y - a]):
        bad.append([a, b, x, y, "unchanged"])
res = [None] * world
dist.all_gather_object(res, {"rank": rank, "rows_checked": int(checked), "bad": bad, "slab_cells": int(slab.rows) * size,
                             "device_bytes": int(slab.field.info()["device_bytes"])})
if rank == 0:
    json.dump({"world": world, "size": size, "cells": size * size, "ranks": res}, open(out, "w"))
dist.destroy_process_group()
