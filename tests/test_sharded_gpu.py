"""The sharded CUDA path.  On one GPU: several slabs of one grid live on the same device and exchange
ghost layers by device copies -- this exercises everything slab-specific in the kernels (ghost offsets,
colour parity of a slab that starts on an odd row, owned-row delta, first/last slab borders) and must be
bit-identical to the reference golden vectors.  With >= 2 GPUs: the real thing over NCCL via torchrun."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import common
from epic_b200 import grids
from epic_b200.sharded import GpuSlab, ShardedSolver

pytestmark = pytest.mark.gpu


class LocalGroup:
    """Drives `world` slabs that live in this process; halo exchange = device-to-device copies."""

    def __init__(self, shape, world, math="strict", p2p=False, tracking=False):
        self.slabs = [GpuSlab(shape, r, world, math=math) for r in range(world)]
        if tracking:    # static-tile skipping, as ShardedSolver.solve switches it on
            for s in self.slabs:
                s.set_tracking(True)
        self.world = world
        self.iteration = 0
        self.delta = 0.0
        self.p2p = p2p
        if p2p:     # the kernels store the edge layers into the neighbour slab themselves
            for r in range(world - 1):
                assert self.slabs[r].field.set_peer_local(1, self.slabs[r + 1].field) == 0
                assert self.slabs[r + 1].field.set_peer_local(0, self.slabs[r].field) == 0

    def upload(self, u, locked):
        for s in self.slabs:
            lo, hi = s.held_range()
            s.upload(u[lo:hi], locked[lo:hi])

    def exchange(self):
        if self.p2p:
            return
        g = self.slabs[0].ghost
        for r in range(self.world - 1):
            a, b = self.slabs[r], self.slabs[r + 1]
            b.halo("recv_up", g).copy_(a.halo("send_down", g))
            a.halo("recv_down", g).copy_(b.halo("send_up", g))

    def run(self, count, check_last=False):
        T = self.slabs[0].T
        done = 0
        while done < count:
            c = min(T, count - done)
            for s in self.slabs:
                s.run_pass(self.iteration + done, c, check_last and done + c == count)
            done += c
            self.exchange()
        self.iteration += count
        if check_last:
            self.delta = max(s.read_delta() for s in self.slabs)

    def run_iterations(self, count, stagger):
        left = count
        while left > 0:
            to_check = (-self.iteration) % stagger
            if to_check < left:
                self.run(to_check + 1, True)
                left -= to_check + 1
            else:
                self.run(left, False)
                left = 0

    def field(self):
        return np.concatenate([s.download_owned() for s in self.slabs], 0)


@pytest.mark.parametrize("tracking", [False, True])
@pytest.mark.parametrize("p2p", [False, True])
@pytest.mark.parametrize("world,case", [(2, "random_ragged"), (3, "random256"), (5, "proc_maze"), (2, "random3d_ragged"),
                                        (3, "random48x3")])
def test_slabs_on_one_gpu_bit_identical(golden, libepic_built, world, case, p2p, tracking):
    u, locked, eps, stagger = common.case_input(case)
    grp = LocalGroup(u.shape, world, p2p=p2p, tracking=tracking)
    grp.upload(u, locked)
    done = 0
    for k in sorted(int(c) for c in golden[case]["checkpoints"]):
        grp.run_iterations(k - done, stagger)
        done = k
        g = golden[case]["checkpoints"][str(k)]
        assert common.sha1(grp.field()) == g["sha1_u"], "%s x%d: field differs after %d iterations" % (case, world, k)
        assert common.hexf(grp.delta) == g["delta_hex"]


def test_world1_driver_on_gpu_solves_golden(golden, libepic_built):
    u, locked, eps, stagger = common.case_input("basic")
    slab = GpuSlab(u.shape, 0, 1)
    slab.upload(u, locked)
    s = ShardedSolver(slab)
    it, delta = s.solve(eps, stagger)
    g = golden["basic"]["complete"]
    assert it == g["iterations"] and common.hexf(delta) == g["delta_hex"]
    assert common.sha1(slab.download_owned()) == g["sha1_u"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_nccl_two_ranks_bit_identical(golden, libepic_built, tmp_path):
    script = os.path.join(common.ROOT, "tests", "sharded_worker.py")
    out = tmp_path / "out.json"
    n = min(torch.cuda.device_count(), 4)
    for halo, case in (("p2p", "proc_maze"), ("nccl", "proc_maze"), ("p2p", "random48x3")):
        subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29731", script, case, str(out), halo],
                       check=True, timeout=600)
        res = json.load(open(out))
        g = golden[case]["complete"]
        assert res["halo"] == halo
        assert res["iterations"] == g["iterations"] and res["delta_hex"] == g["delta_hex"]
        assert res["sha1_u"] == g["sha1_u"], "%s over %s differs from the reference" % (case, halo)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_maximum_size_2d_sharded_bands_against_oracle(libepic_built, tmp_path):
    """BASELINE.json config 4's size (65536^2 = 2^32 cells, 33 GiB resident) row-sharded over every GPU of the box:
    see tests/maxsize_worker.py."""
    out = tmp_path / "maxsize.json"
    n = min(torch.cuda.device_count(), 8)
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                    "--master-addr", "127.0.0.1", "--master-port", "29733", os.path.join(common.ROOT, "tests", "maxsize_worker.py"),
                    "65536", str(out)], check=True, timeout=900)
    res = json.load(open(out))
    assert res["cells"] == 2 ** 32 and res["world"] == n
    for r in res["ranks"]:
        assert r["bad"] == [], "rank %d: bands differ from the 64-bit oracle: %r" % (r["rank"], r["bad"])
        assert r["rows_checked"] >= 96
    assert sum(r["rows_checked"] for r in res["ranks"]) == 2 * (192) - 2 * 12 + (n - 1) * (192 - 24)


@pytest.mark.parametrize("p2p", [False, True])
def test_sharded_solve_with_static_tile_skipping_equals_one_field(libepic_built, p2p):
    """Solve to epsilon on three slabs with static-tile skipping on (edge tiles always run): iteration count,
    delta and field must equal the single-field solve (itself pinned to the golden vectors), and interior
    tiles must actually have been skipped."""
    from epic_b200.field import Field
    shape = (1500, 1024)
    u, locked = grids.random_obstacles(shape, 0.2, 3, seed=31)
    f = Field(shape)
    f.upload(u, locked)
    it1, d1 = f.solve(1e-3, 100)
    want = common.sha1(f.download_u())
    f.close()
    grp = LocalGroup(shape, 3, p2p=p2p, tracking=True)
    grp.upload(u, locked)
    while True:
        grp.run((-grp.iteration) % 100 + 1, True)
        if grp.delta < 1e-3 and grp.iteration >= max(shape):
            break
    assert (grp.iteration, grp.delta) == (it1, d1)
    assert common.sha1(grp.field()) == want
    assert sum(s.field.info()["skipped_tiles"] for s in grp.slabs) > 0
