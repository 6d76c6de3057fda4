#!/usr/bin/env python
"""x0-slab sharded 3-D sweep throughput (BASELINE.json config 5) under torch.distributed.run: every rank holds
m0/W layers of an m^3 random-obstacle grid, halos by peer-to-peer stores, one all-reduce(max) per check.
usage: python -m torch.distributed.run --nproc-per-node W tools/sharded3d_timing.py [size] [strict|fast] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from epic_b200 import grids  # noqa: E402
from epic_b200.sharded import GpuSlab, ShardedSolver  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
math = sys.argv[2] if len(sys.argv) > 2 else "fast"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
shape = (size, size, size)
slab = GpuSlab(shape, rank, world, math=math)
lo, hi = slab.held_range()
u, locked = grids.random_obstacles(shape, 0.2, 64, seed=1234, row0=lo, rows=hi - lo)
slab.upload(u, locked)
del u, locked
solver = ShardedSolver(slab)
solver.run(1, True)
solver.run(100, True)
torch.cuda.synchronize()
dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    solver.run(100, True)
b.record()
torch.cuda.synchronize()
dist.barrier()
t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item())
    print("3-D %d^3 %s on %d GPUs (x0 slabs, p2p halos): %.2f ms per 100 half-sweeps, %.1f Gcell-updates/s, delta %.6g" % (
        size, math, world, ms / steps, size ** 3 / 2 * 100 * steps / (ms * 1e-3) / 1e9, solver.delta), flush=True)
dist.destroy_process_group()
