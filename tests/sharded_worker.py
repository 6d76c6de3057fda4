"""torchrun worker of tests/test_sharded_gpu.py: solve a golden case sharded over all ranks (NCCL)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import common  # noqa: E402
from epic_b200.sharded import GpuSlab, ShardedSolver, gather_field  # noqa: E402

case, out = sys.argv[1], sys.argv[2]
halo = sys.argv[3] if len(sys.argv) > 3 else "p2p"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
u, locked, eps, stagger = common.case_input(case)
slab = GpuSlab(u.shape, rank, world, halo=halo)
lo, hi = slab.held_range()
slab.upload(u[lo:hi], locked[lo:hi])
solver = ShardedSolver(slab)
it, delta = solver.solve(eps, stagger)
field = gather_field(slab)
if rank == 0:
    json.dump({"iterations": it, "delta_hex": common.hexf(delta), "sha1_u": common.sha1(field),
               "exchanges": solver.exchanges, "halo": "p2p" if slab.p2p else "nccl"}, open(out, "w"))
dist.destroy_process_group()
