"""Host-side logic of the multi-GPU path, on the CPU: world_size-2 and -3 process groups over gloo run
epic_b200.sharded.ShardedSolver (the product's partitioning, pass/exchange schedule, all-reduce(max) and
termination rule) on a test double of the slab that sweeps with the oracle instead of CUDA.  The sharded
result must be bit-identical to the unsharded oracle, and to the reference golden vectors.

The GPU twin (same driver on GpuSlab over NCCL) is tests/test_sharded_gpu.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
from epic_b200 import grids
from epic_b200.sharded import ShardedSolver, gather_field, partition
from oracle import oracle as orc


class OracleSlab:
    """Same interface as epic_b200.sharded.GpuSlab; the sweeps are the oracle's (test double)."""

    def __init__(self, shape, rank, world, T):
        self.shape = tuple(shape)
        self.rank, self.world = rank, world
        self.row0, self.rows = partition(self.shape[0], world, rank)
        self.T = T
        self.ghost = T if world > 1 else 0
        self._delta = 0.0
        self.nlaunch = 0

    def held_range(self):
        return max(0, self.row0 - self.ghost), min(self.shape[0], self.row0 + self.rows + self.ghost)

    def upload(self, u, locked):
        lo, hi = self.held_range()
        assert u.shape[0] == hi - lo
        # storage always has ghost layers on both sides (like the device buffers), even outside the grid
        full = (self.rows + 2 * self.ghost,) + self.shape[1:]
        self.u = np.full(full, -1e6, np.float32)
        self.locked = np.ones(full, np.uint32)
        off = lo - (self.row0 - self.ghost)
        self.u[off:off + hi - lo] = u
        self.locked[off:off + hi - lo] = locked
        self.lo_off, self.hi_off = off, off + hi - lo

    def download_owned(self):
        return self.u[self.ghost:self.ghost + self.rows].copy()

    def run_pass(self, it0, count, check_last):
        assert count <= self.T or self.world == 1   # a lone slab has no ghost layers to go stale
        self.nlaunch += 1
        view_u = self.u[self.lo_off:self.hi_off]
        view_l = self.locked[self.lo_off:self.hi_off]
        g0 = self.row0 - self.ghost + self.lo_off            # global layer of view row 0
        o = orc.Oracle(view_u, view_l, 1e-3, 1 << 30)
        assert o.u is view_u or np.shares_memory(o.u, view_u)
        own = slice(self.row0 - g0, self.row0 - g0 + self.rows)
        for s in range(count):
            before = view_u[own].copy() if (check_last and s == count - 1) else None
            # the oracle's colour rule uses its local layer index: shift the iteration parity by g0
            o.iteration = it0 + s + (g0 & 1)
            o.update()
            if before is not None:
                self._delta = float(np.abs(before - view_u[own]).max())

    def read_delta(self):
        d, self._delta = self._delta, 0.0
        return d

    def halo(self, which, layers):
        first = {"send_up": self.ghost, "send_down": self.ghost + self.rows - layers,
                 "recv_up": self.ghost - layers, "recv_down": self.ghost + self.rows}[which]
        return torch.from_numpy(self.u[first:first + layers])

    def scalar(self, value):
        return torch.tensor([value], dtype=torch.float32)

    def launches(self):
        return self.nlaunch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case, T, iterations, solve, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        u, locked, eps, stagger = common.case_input(case)
        slab = OracleSlab(u.shape, rank, world, T)
        lo, hi = slab.held_range()
        slab.upload(u[lo:hi], locked[lo:hi])
        solver = ShardedSolver(slab)
        if solve:
            solver.solve(eps, stagger)
        else:
            solver.run_iterations(iterations, stagger)
        field = gather_field(slab)
        if rank == 0:
            out.put((solver.iteration, solver.delta, common.sha1(field), solver.exchanges))
    finally:
        dist.destroy_process_group()


def run_sharded(world, case, T, iterations=0, solve=False):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, T, iterations, solve, out)) for r in range(world)]
    for p in procs:
        p.start()
    result = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return result


def test_partition_covers_grid():
    for m0 in (5, 64, 131, 16384, 65536):
        for world in (1, 2, 3, 8):
            spans = [partition(m0, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(s[1] for s in spans) == m0
            for a, b in zip(spans, spans[1:]):
                assert a[0] + a[1] == b[0]
            assert max(s[1] for s in spans) - min(s[1] for s in spans) <= 1


def test_single_rank_driver_matches_golden(golden):
    """world = 1: the driver's schedule (passes of T, checks closing a pass) against the reference vectors."""
    u, locked, eps, stagger = common.case_input("random_ragged")
    slab = OracleSlab(u.shape, 0, 1, 4)
    slab.upload(u, locked)
    s = ShardedSolver(slab)
    s.solve(eps, stagger)
    g = golden["random_ragged"]["complete"]
    assert s.iteration == g["iterations"] and common.hexf(s.delta) == g["delta_hex"]
    assert common.sha1(slab.download_owned()) == g["sha1_u"]


@pytest.mark.parametrize("world,case,T", [(2, "random_ragged", 4), (3, "random256", 4), (2, "random3d_ragged", 1)])
def test_sharded_fixed_iterations_bit_identical(golden, world, case, T):
    k = max(int(c) for c in golden[case]["checkpoints"])
    it, delta, sha, exchanges = run_sharded(world, case, T, iterations=k)
    g = golden[case]["checkpoints"][str(k)]
    assert it == k and sha == g["sha1_u"] and common.hexf(delta) == g["delta_hex"]
    assert exchanges >= k // T


@pytest.mark.parametrize("world,case,T", [(2, "box64", 4), (2, "random3d_ragged", 1)])
def test_sharded_solve_to_epsilon_bit_identical(golden, world, case, T):
    it, delta, sha, _ = run_sharded(world, case, T, solve=True)
    g = golden[case]["complete"]
    assert it == g["iterations"] and common.hexf(delta) == g["delta_hex"] and sha == g["sha1_u"]
