"""Row-slab (2-D) / x0-slab (3-D) sharding of one harmonic grid over several B200s, one process per GPU.

The reference is single-device (no collectives anywhere); this is the multi-GPU
step BASELINE.json's north_star adds.  Rank r owns the x0-layers [r*m0/W, (r+1)*m0/W) plus `ghost`
layers on each side that mirror the neighbours' edge layers.  A pass (up to T half-sweeps fused in one
kernel) needs ghost layers that were correct at its start, so after every pass each rank's first / last
T owned layers must reach the rank above / below.  Two transports:
  halo="p2p"   (default on one node) the sweep kernel itself stores those layers into the neighbours'
               ghost layers over NVLink (CUDA IPC peer mappings) and passes are ordered between GPUs by
               stream-ordered flag writes / waits: nothing on the host and no collective per pass;
  halo="nccl"  NCCL send/recv through torch.distributed after every pass (2 messages of T*pitch floats
               per neighbour).
Every check sweep is followed by ONE all-reduce(max) of the per-rank delta.  Red-black ordering makes the
result independent of the partition: the sharded field is bit-identical to the single-GPU (and the CPU)
field.

The driver is written against a small slab interface so that the same exchange schedule runs
  * on the product slab (epic_b200.field.Field, CUDA) -- GpuSlab below, and
  * in the CPU test-suite on a test double that sweeps with the oracle over gloo (tests/test_sharded.py).
"""
import numpy as np
import torch
import torch.distributed as dist

from .field import DevicePointer, Field


def partition(m0, world, rank):
    """Owned x0-range of `rank`: contiguous, sizes differ by at most one layer."""
    lo = (m0 * rank) // world
    hi = (m0 * (rank + 1)) // world
    return lo, hi - lo


class GpuSlab:
    """One rank's slab on its GPU, running on torch's current CUDA stream so that kernels and the
    NCCL transfers torch issues are ordered on the device without host synchronisation."""

    def __init__(self, shape, rank, world, math="strict", device=None, halo="p2p", group=None):
        self.shape = tuple(int(s) for s in shape)
        self.rank, self.world = rank, world
        self.row0, self.rows = partition(self.shape[0], world, rank)
        dev = torch.cuda.current_device() if device is None else device
        stream = torch.cuda.current_stream(dev).cuda_stream
        # ghost depth = sweeps per pass; query T from a probe of the library defaults
        probe_T = 4 if len(self.shape) == 2 else 2
        self.ghost = probe_T if world > 1 else 0
        self.field = Field(self.shape, self.row0, self.rows, self.ghost, math=math, device=dev, stream=stream)
        info = self.field.info()
        self.T = info["sweeps_per_pass"]
        assert world == 1 or self.T == self.ghost
        self.layer_floats = info["layer_floats"]
        self._views = {}
        self.group = group
        self.p2p = False
        if world > 1 and halo == "p2p" and dist.is_initialized():
            self._connect_peers()

    def _connect_peers(self):
        """Swap CUDA IPC handles with the neighbouring ranks; all ranks agree on the outcome."""
        blobs = [None] * self.world
        dist.all_gather_object(blobs, self.field.peer_export(), group=self.group)
        ok = 1
        if self.rank > 0 and self.field.set_peer_ipc(0, blobs[self.rank - 1]) != 0:
            ok = 0
        if self.rank < self.world - 1 and self.field.set_peer_ipc(1, blobs[self.rank + 1]) != 0:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:
            raise RuntimeError("peer-to-peer halo setup failed on some rank (CUDA IPC / peer access); "
                               "construct GpuSlab(..., halo='nccl') to use NCCL send/recv instead")
        self.p2p = True

    # global layers held (owned + ghost), clipped to the grid
    def held_range(self):
        lo = max(0, self.row0 - self.ghost)
        hi = min(self.shape[0], self.row0 + self.rows + self.ghost)
        return lo, hi

    def upload(self, u, locked):
        """u, locked: dense arrays covering held_range()."""
        lo, hi = self.held_range()
        self.field.upload(u, locked, first=lo, layers=hi - lo)
        if self.p2p:
            # a neighbour must not start storing into this slab's ghost layers before the upload is done
            torch.cuda.synchronize()
            dist.barrier(group=self.group)

    def download_owned(self):
        return self.field.download_u(first=self.row0, layers=self.rows)

    def run_pass(self, it0, count, check_last):
        self.field.run(it0, count, check_last)

    def set_tracking(self, on):
        """Static-tile skipping (bit-identical, see include/epic_b200.h) for the passes that follow."""
        self.field.set_tracking(on)

    def read_delta(self):
        return self.field.read_delta()

    def _view(self, layer, layers):
        ptr = self.field.layer_ptr(layer)
        key = (ptr, layers)
        if key not in self._views:
            self._views[key] = torch.as_tensor(DevicePointer(ptr, (layers * self.layer_floats,)), device="cuda")
        return self._views[key]

    def halo(self, which, layers):
        """Tensor views of the CURRENT buffer: 'send_up' = my first owned layers, 'send_down' = my last
        owned layers, 'recv_up' = ghost above, 'recv_down' = ghost below."""
        first = {"send_up": self.row0, "send_down": self.row0 + self.rows - layers,
                 "recv_up": self.row0 - layers, "recv_down": self.row0 + self.rows}[which]
        return self._view(first, layers)

    def scalar(self, value):
        return torch.tensor([value], dtype=torch.float32, device="cuda")

    def launches(self):
        return self.field.info()["launches"]

    def sync(self):
        self.field.sync()


class ShardedSolver:
    """The reference's update / update_and_check / execute loop (libepic/src/harmonic/harmonic_gpu.cu:226-415)
    over a sharded grid.  Every rank runs the same calls; `iteration` and `delta` are global values."""

    def __init__(self, slab, group=None):
        self.slab = slab
        self.group = group
        self.rank, self.world = slab.rank, slab.world
        self.iteration = 0
        self.delta = 0.0
        self.exchanges = 0

    def exchange(self):
        """Refresh the ghost layers from the neighbours' freshly written edge layers."""
        if self.world == 1 or getattr(self.slab, "p2p", False):
            return          # p2p: the sweep kernel has already stored the edge layers into the neighbours
        g = self.slab.ghost
        ops = []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, self.slab.halo("send_up", g), self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, self.slab.halo("recv_up", g), self.rank - 1, self.group))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, self.slab.halo("send_down", g), self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, self.slab.halo("recv_down", g), self.rank + 1, self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()     # on NCCL this orders the current stream after the transfer; no host sync
        self.exchanges += 1

    def run(self, count, check_last=False):
        """`count` half-sweeps from self.iteration; passes of T with a halo exchange after each."""
        T = self.slab.T
        done = 0
        if self.world == 1 or getattr(self.slab, "p2p", False):
            # nothing to do between passes on the host: let the library cut the range into passes
            self.slab.run_pass(self.iteration, count, check_last)
            done = count
        while done < count:
            c = min(T, count - done)
            self.slab.run_pass(self.iteration + done, c, check_last and done + c == count)
            done += c
            self.exchange()
        self.iteration += count
        if check_last:
            local = self.slab.read_delta()
            if self.world > 1:
                t = self.slab.scalar(local)
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
                local = float(t.item())
            self.delta = local
        return self.delta

    def update(self):
        self.run(1, False)
        return 0

    def update_and_check(self, epsilon):
        self.run(1, True)
        return 1 if self.delta < epsilon else 0

    def run_iterations(self, count, stagger):
        """The complete() schedule without the termination test, fused into passes between checks."""
        left = count
        while left > 0:
            # next check sweep is the first iteration >= current that is a multiple of stagger
            to_check = (-self.iteration) % stagger
            if to_check < left:
                self.run(to_check + 1, True)
                left -= to_check + 1
            else:
                self.run(left, False)
                left = 0

    def solve(self, epsilon, stagger, m_max=None, max_iterations=None):
        """harmonic_execute_gpu's loop: stop right after a check sweep with delta < epsilon once
        iteration >= max(m).  Returns (iterations, delta).  `max_iterations` (not in the reference, whose loop
        is unbounded) raises TimeoutError once that many iterations have run without convergence."""
        if not epsilon > 0.0 or stagger <= 0:
            raise ValueError("epsilon must be positive and stagger non-zero")
        m_max = max(self.slab.shape) if m_max is None else m_max
        self.iteration = 0
        if hasattr(self.slab, "set_tracking"):
            self.slab.set_tracking(True)     # tiles that stopped changing (and do not read ghost layers) are skipped
        try:
            while True:
                to_check = (-self.iteration) % stagger
                self.run(to_check + 1, True)
                if self.delta < epsilon and self.iteration >= m_max:
                    return self.iteration, self.delta
                if max_iterations is not None and self.iteration >= max_iterations:
                    raise TimeoutError("no convergence after %d iterations (delta %g)" % (self.iteration, self.delta))
        finally:
            if hasattr(self.slab, "set_tracking"):
                self.slab.set_tracking(False)


def gather_field(slab, group=None):
    """The whole field on rank 0 (numpy), None elsewhere: the sharded twin of
    harmonic_get_potential_values_gpu."""
    own = torch.from_numpy(np.ascontiguousarray(slab.download_owned()))
    if slab.world == 1:
        return own.numpy()
    backend = dist.get_backend(group)
    if backend == "nccl":
        own = own.cuda()
    sizes = [partition(slab.shape[0], slab.world, r)[1] for r in range(slab.world)]
    parts = [torch.empty((s,) + tuple(own.shape[1:]), dtype=own.dtype, device=own.device) for s in sizes]
    dist.all_gather(parts, own, group=group) if len(set(sizes)) == 1 else _all_gather_ragged(parts, own, slab, group)
    if slab.rank != 0:
        return None
    return torch.cat(parts, 0).cpu().numpy()


def _all_gather_ragged(parts, own, slab, group):
    for r in range(slab.world):
        if r == slab.rank:
            parts[r].copy_(own)
        dist.broadcast(parts[r], src=r, group=group)
