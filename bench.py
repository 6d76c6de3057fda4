#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: Gcell-updates/s (and time-to-epsilon) of the log-space harmonic
relaxation on the synthetic 16384 x 16384 random-obstacle grid (configs[2]), at 1/2/4/8 B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

A "step" is one stagger period of the reference's solver loop over the whole grid: 100 red-black
half-sweeps, the last one a convergence-check sweep (reference libepic/src/harmonic/harmonic_gpu.cu:266-290),
i.e. 100 * (16384*16384/2) lattice-site updates.  One lattice-site update = one cell of the active colour in
one half-sweep; algorithmic traffic 8 B per update.

  value     device-resident throughput: grid already in HBM, K steps timed with CUDA events, max over ranks
  e2e       the same steps through the reference-facing call path with HOST buffers inside the timed
            region: harmonic_update_model_gpu (H2D of u and locked) -> harmonic_update_and_check_gpu +
            99 x harmonic_update_gpu -> harmonic_get_potential_values_gpu (D2H of u)   [N = 1: libepic C ABI;
            N > 1: the slab API, each rank moving its own slab]
  roofline  HBM roofline of the sweep kernel: 8 B x updates / kernel time vs MEASURED_PEAKS.json
  cpu_baseline  the reference's own CPU code (oracle/_ref, built from the untouched sources) on this host
  gpu_baseline  the reference's own GPU code (harmonic_gpu.cu recompiled for sm_100a, oracle/_ref) on this B200:
                harmonic_update_gpu x K on the same grid and harmonic_complete_gpu on maps/maze.png, beside this
                library's figures for the same calls (N = 1 only; a subprocess under a timeout)
  tte_* / field_sha1*  time to epsilon (every tile swept / static tiles skipped), its iteration count, and the sha1
                of the converged field gathered from all ranks -- compared with the hash committed for N = 1
                (tests/golden/bench_fields.json), so that a scaling run carries its own correctness evidence
  abi_multi     (N > 1, rank 0 after the ranks have finished) the same grid through the libepic C ABI alone, with
                EPIC_DEVICES naming all N GPUs: the single-process sharding inside the library (engine/grid.cu)
The grid (1 GiB of potentials per buffer) is 8x larger than L2, so nothing survives between passes.

  --workload maze --size 65536 --gpus 8   BASELINE.json config 4 (procedural maze, 16 GiB field)
  --dims 3 --size 1024 [--gpus N]         BASELINE.json config 5
"""
import hashlib
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SWEEPS_PER_STEP = 100
ALGO_BYTES_PER_UPDATE = 8.0


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled while the timed region runs: NVML polled every few milliseconds
    from a thread (the timed region of a multi-GPU run is only tens of milliseconds long), nvidia-smi as the
    fallback when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.sm, self.max_sm, self.reasons = [], None, set()
        self.nvml, self.handle, self.stop_flag, self.thread = None, None, False, None

    def _poll(self):
        n = self.nvml
        names = (("hw_slowdown", getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                 ("hw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                 ("sw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                 ("sw_power_cap", getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)))
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                mask = int(get_reasons(self.handle))
                for name, bit in names:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def mark(self):
        """Forget what was sampled so far (called right before the timed region starts)."""
        self.sm, self.rows, self.reasons = [], [], set()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        if self.thread is not None:
            self.thread.join(timeout=2)

    def summary(self):
        if self.nvml is not None:
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, 4 ms poll"}
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def grid_shape(args):
    return (args.size,) * args.dims


def make_grid(args, row0=0, rows=None):
    """(u, locked) of rows [row0, row0 + rows) of the bench grid."""
    from epic_b200 import grids
    shape = grid_shape(args)
    if args.workload == "maze":
        return grids.procedural_maze(shape, corridor=args.corridor, wall=2, goals=args.goals, seed=1234, row0=row0, rows=rows)
    return grids.random_obstacles(shape, 0.2, args.goals, seed=1234, row0=row0, rows=rows)


def workload_name(args):
    shape = "x".join(str(s) for s in grid_shape(args))
    if args.workload == "maze":
        return "synthetic %s procedural maze (corridor %d, wall 2, %d goals, seed 1234)" % (shape, args.corridor, args.goals)
    return "synthetic %s random-obstacle grid (p=0.2, %d goals, seed 1234)" % (shape, args.goals)


def workload(args, sweeps_per_step=None):
    shape = grid_shape(args)
    cells = int(np.prod(shape, dtype=np.int64))
    if sweeps_per_step is not None:     # the reference arm: its step is one half-sweep of a band
        return {"workload": workload_name(args) + ", %s-slab sharded" % ("row" if args.dims == 2 else "x0"),
                "grid": list(shape), "epsilon": 1e-3, "stagger": SWEEPS_PER_STEP, "sweeps_per_step": sweeps_per_step,
                "parallelism": "host cores"}
    return {"workload": workload_name(args) + ", %s-slab sharded" % ("row" if args.dims == 2 else "x0"),
            "grid": list(shape), "epsilon": 1e-3, "stagger": SWEEPS_PER_STEP,
            "sweeps_per_step": SWEEPS_PER_STEP, "updates_per_step": cells // 2 * SWEEPS_PER_STEP,
            "parallelism": "%s-slab x%d" % ("row" if args.dims == 2 else "x0", args.gpus),
            "halo": (args.halo if args.gpus > 1 else None),
            "l2": "grid (%.2f GiB per buffer) is larger than L2; no flush needed" % (cells * 4 / 2**30)}


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code on the host cores

def cpu_reference_rate(args, rows, sweeps, warmup=0):
    """(updates/s, description) of the reference CPU half-sweep over the first `rows` x0-layers of the grid."""
    from oracle import oracle as orc
    size, dims = args.size, args.dims
    u, locked = make_grid(args, 0, rows)
    locked[-1] = 1
    if orc.have_ref():
        solver, kind, cores = orc.Reference(u, locked, 1e-3, SWEEPS_PER_STEP), "reference", 1
        what = "harmonic_update_cpu of the untouched reference sources (oracle/_ref), serial as shipped"
    else:
        cores = os.cpu_count() or 1
        solver, kind = orc.Oracle(u, locked, 1e-3, SWEEPS_PER_STEP, threads=cores), "port"
        what = "oracle port (oracle/harmonic_oracle.c), OpenMP over rows"
    solver.iteration = 1
    for _ in range(warmup):
        solver.update()
    t0 = time.perf_counter()
    for _ in range(sweeps):
        solver.update()
    dt = time.perf_counter() - t0
    updates = (rows - 2) * float(size - 2) ** (dims - 1) / 2.0 * sweeps
    return updates / dt, dt, kind, cores, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # one step = one half-sweep over a band sized so that the whole run stays within ~150 s at ~25 M updates/s
    budget_updates = 150.0 * 25e6 / max(1, args.steps + args.warmup)
    rows = int(min(args.size, max(8, budget_updates / (float(args.size) ** (args.dims - 1) / 2.0))))
    rate, dt, kind, cores, what = cpu_reference_rate(args, rows, args.steps, args.warmup)
    value = rate / 1e9
    sample = "%d half-sweeps (steps) over x0 = 0..%d of the %s grid; %s" % (
        args.steps, rows - 1, "x".join(str(v) for v in grid_shape(args)), what)
    line = {"impl": "reference", "metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload(args, sweeps_per_step=1), step="one half-sweep (harmonic_update_cpu) over x0 = 0..%d; "
                           "the rate is per lattice-site update, the unit the native arm reports" % (rows - 1),
                           updates_per_step=int((rows - 2) * float(args.size - 2) ** (args.dims - 1) / 2.0)),
            "gpu_launches": 0,
            "cpu_baseline": {"value": value, "unit": "Gcell-updates/s", "cores": cores, "kind": kind, "sample": sample,
                             "host_cores": os.cpu_count()},
            "e2e": {"value": value, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def hash_band(ndim):
    """x0-layers per band of the field hash: slab boundaries of 1/2/4/8 ranks must fall on band boundaries
    (16384 / 8 = 2048 rows, 65536 / 8 = 8192 rows; 1024 / 8 = 128 layers, 512 / 8 = 64 layers)."""
    return 1024 if ndim == 2 else 32


def band_digests(own, row0):
    """sha1 digests of the bands of this rank's owned layers (row0 must be a band boundary unless the rank owns
    everything)."""
    out, band = [], hash_band(own.ndim)
    for a in range(0, own.shape[0], band):
        out.append(((row0 + a) // band, hashlib.sha1(np.ascontiguousarray(own[a:a + band]).tobytes()).hexdigest()))
    return out


def field_hash(digests):
    """One hash of the whole field from the per-band digests of all ranks (independent of the partition)."""
    h = hashlib.sha1()
    for _, d in sorted(digests):
        h.update(bytes.fromhex(d))
    return h.hexdigest()


def committed_fields():
    path = os.path.join(ROOT, "tests", "golden", "bench_fields.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


def source_stamp():
    """Hash of the sweep-kernel sources: profiles/ncu_traffic.json carries the stamp of the kernels it was measured
    on, and a stale measurement is reported as null instead of silently describing another kernel."""
    h = hashlib.sha1()
    for name in ("sweep2d.cuh", "sweep3d.cuh", "math_policies.cuh", "strict_math.h"):
        with open(os.path.join(ROOT, "epic_b200", "csrc", "kernels", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def reference_gpu_baseline(args):
    """The reference's own GPU path (recompiled for sm_100a) on this box, in a subprocess under a timeout, and this
    library's time for the same harmonic_complete_gpu call on maps/maze.png."""
    from oracle import ref_gpu
    if not ref_gpu.available():
        return {"unavailable": "oracle/_ref/libepic_ref_gpu.so was not built (no /root/reference at build time)"}
    out = {}

    def sub(argv, timeout):
        try:
            r = subprocess.run([sys.executable, "-m", "oracle.ref_gpu"] + argv, cwd=ROOT, capture_output=True, text=True,
                               timeout=timeout)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            return json.loads(lines[-1]) if lines else {"error": (r.stderr or "no output")[-300:]}
        except subprocess.TimeoutExpired:
            return {"error": "timed out after %d s" % timeout}
    # The stock build: its sweep kernels call __syncthreads() in divergent code (reference harmonic_gpu.cu:46-49,
    # :86-89), which deadlocks on Volta and later (SASS: WARPSYNC.ALL on both sides of the divergent branch;
    # profiles/r02_reference_gpu.md).  The numbers come from the build with the barriers compiled out
    # (oracle/Makefile refgpu_nobar), through harmonic_update_gpu, which needs none.
    if args.try_stock_reference_gpu:
        out["stock_complete_gpu_maze"] = sub(["complete", "--map", "maze"], 25)
    else:
        out["stock_complete_gpu_maze"] = {
            "skipped": "the stock build deadlocks on sm_100a: killed after 25 s in profiles/r02c_bench_default.json, after 300 / "
                       "300 / 600 s in profiles/r02_reference_gpu.md; --try-stock-reference-gpu repeats the 25-second attempt "
                       "(left out of the default run so that a wedged kernel cannot disturb what runs on the box afterwards)"}
    if ref_gpu.available("nobar"):
        if args.dims == 2 and args.size <= 16384:
            out["sweeps"] = sub(["sweeps", "--variant", "nobar", "--size", str(args.size), "--steps", "10", "--warmup", "2"], 300)
        out["maze"] = sub(["updates", "--variant", "nobar", "--map", "maze", "--iterations", "49301"], 180)
    # ours, same call, same map
    try:
        from epic_b200 import grids
        from epic_b200.harmonic import Harmonic
        maps = np.load(os.path.join(ROOT, "tests", "golden", "maps.npz"))
        u, locked = grids.grid_from_image(maps["maze"])
        best = None
        for _ in range(3):
            h = Harmonic(u.copy(), locked.copy(), 1e-3, 100)
            t0 = time.perf_counter()
            h.solve(process="gpu")
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out["maze_ours"] = {"seconds": best, "iterations": int(h.currentIteration), "delta": float(h.delta),
                            "what": "harmonic_complete_gpu of this library on the same map (strict unless EPIC_MATH says otherwise), best of 3"}
    except Exception as e:      # noqa: BLE001
        out["maze_ours"] = {"error": str(e)[:200]}
    return out


# ---------------------------------------------------------------------------------------------------
# native arm

def run_native(args):
    import torch
    import torch.distributed as dist

    from epic_b200 import grids, libepic
    from epic_b200.sharded import GpuSlab, ShardedSolver, partition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run --nproc-per-node %d"
                         % (args.gpus, world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = libepic.load()
    size = args.size
    shape = grid_shape(args)
    layer_cells = size ** (args.dims - 1)
    updates_per_step = size * layer_cells // 2 * SWEEPS_PER_STEP

    # what this rank holds (owned rows + ghost rows), generated once and kept in pinned host memory: the
    # e2e leg moves it inside the timed region
    from epic_b200.sharded import partition as _partition
    row0, nrows = _partition(size, world, rank)
    ghost = (4 if args.dims == 2 else 2) if world > 1 else 0
    lo, hi = max(0, row0 - ghost), min(size, row0 + nrows + ghost)
    u_held, locked_held = make_grid(args, lo, hi - lo)
    u_pin = torch.from_numpy(u_held).pin_memory()
    l_pin = torch.from_numpy(locked_held.view(np.int32)).pin_memory()
    u_host, l_host = u_pin.numpy(), l_pin.numpy().view(np.uint32)
    del u_held, locked_held

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_src = peaks()

    def measure(math, with_tte):
        """All numbers of one arithmetic mode."""
        os.environ["EPIC_MATH"] = math      # the libepic C ABI takes its options from the environment
        slab = GpuSlab(shape, rank, world, math=math, halo=args.halo)
        assert slab.held_range() == (lo, hi)
        slab.upload(u_host, l_host)
        solver = ShardedSolver(slab)

        # ---- device-resident throughput ----
        with ClockSampler(local) as clocks:
            solver.run(1, True)              # iteration 0 (a check sweep); steps then cover 1..100, 101..200, ...
            for _ in range(args.warmup):
                solver.run(SWEEPS_PER_STEP, True)
            launches0 = slab.launches()
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            clocks.mark()
            start.record()
            for _ in range(args.steps):
                solver.run(SWEEPS_PER_STEP, True)
            stop.record()
            barrier()
            clocks.stop_flag = True
        ms = max_over_ranks(start.elapsed_time(stop))
        launches = slab.launches() - launches0
        value = updates_per_step * args.steps / (ms * 1e-3) / 1e9
        delta_after = solver.delta
        info = slab.field.info()

        # ---- kernel-only timing for the roofline: the pass kernel alone, this rank's share ----
        passes = 50
        barrier()
        start.record()
        for i in range(passes):
            slab.run_pass(solver.iteration + info["sweeps_per_pass"] * i, info["sweeps_per_pass"], False)
        stop.record()
        torch.cuda.synchronize()
        kern_ms = start.elapsed_time(stop) / passes
        own_updates_per_pass = slab.rows * layer_cells // 2 * info["sweeps_per_pass"]
        achieved = own_updates_per_pass * ALGO_BYTES_PER_UPDATE / (kern_ms * 1e-3) / 1e9
        traffic, traffic_note = None, "no ncu measurement for this configuration"
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath) and size == 16384 and args.dims == 2 and args.workload == "random":
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("source_stamp") == source_stamp():
                traffic = tj.get(math)
                traffic = traffic / world if traffic else None
                traffic_note = tj.get("note")
            else:
                traffic_note = "profiles/ncu_traffic.json was measured on other kernel sources (stamp %s, now %s): not reported" % (
                    tj.get("source_stamp"), source_stamp())
        # the throw-away passes above ran without halo exchange: restore a consistent state
        slab.upload(u_host, l_host)
        solver.iteration = 0

        # ---- end to end, host buffers in the timed region ----
        e2e_steps = max(1, min(args.steps, 5))
        if world == 1:
            from epic_b200.harmonic import Harmonic
            h = Harmonic(u_host, l_host, 1e-3, SWEEPS_PER_STEP)
            h.initialize_gpu()
            h2d = u_host.nbytes + l_host.nbytes
            d2h = u_host.nbytes

            def e2e_step():
                h.currentIteration = 0
                h.update_model_gpu()
                h.run_iterations(SWEEPS_PER_STEP, "gpu")
                h.get_potential_values_gpu()
        else:
            out_pin = torch.empty((slab.rows,) + shape[1:], dtype=torch.float32).pin_memory()
            h2d = u_host.nbytes + l_host.nbytes
            d2h = out_pin.numel() * 4

            def e2e_step():
                slab.upload(u_host, l_host)
                solver.iteration = 0
                solver.run(1, True)
                solver.run(SWEEPS_PER_STEP - 1, False)
                slab.field.download_u(first=slab.row0, layers=slab.rows, out=out_pin.numpy())
        u_keep = u_host.copy() if world == 1 else None     # the C ABI downloads into the caller's u array
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e_value = updates_per_step * e2e_steps / (e2e_ms * 1e-3) / 1e9
        if world == 1:
            h.uninitialize_gpu()
            u_host[:] = u_keep
            del u_keep

        # ---- time to epsilon (reported beside the throughput; one run, not part of the timed steps) ----
        tte = None
        if with_tte:
            def run_to_epsilon(native_loop):
                slab.upload(u_host, l_host)
                solver.iteration = 0
                skipped0 = slab.field.info()["skipped_tiles"]
                barrier()
                t0 = time.perf_counter()
                if native_loop and world == 1:   # harmonic_execute_gpu's loop inside the library (Field::solve)
                    it, delta = slab.field.solve(1e-3, SWEEPS_PER_STEP, size)
                    converged = True
                elif native_loop:                # the sharded driver's loop, static-tile skipping switched on
                    try:
                        it, delta = solver.solve(1e-3, SWEEPS_PER_STEP, size, max_iterations=args.tte_max_iterations)
                        converged = True
                    except TimeoutError:
                        it, delta, converged = solver.iteration, solver.delta, False
                else:                # the same loop driven from here, pass by pass, every tile swept
                    converged = False
                    while solver.iteration < args.tte_max_iterations:
                        solver.run((-solver.iteration) % SWEEPS_PER_STEP + 1, True)
                        if solver.delta < 1e-3 and solver.iteration >= size:
                            converged = True
                            break
                    it, delta = solver.iteration, solver.delta
                barrier()
                seconds = max_over_ranks((time.perf_counter() - t0) * 1e3) / 1e3
                return seconds, it, delta, converged, slab.field.info()["skipped_tiles"] - skipped0

            if args.tte_all_tiles:
                sec_all, it_all, delta_all, conv_all, _ = run_to_epsilon(False)
            else:
                sec_all, it_all, delta_all, conv_all = None, None, None, None
            tte = {"seconds": sec_all, "iterations": it_all, "delta": delta_all, "epsilon": 1e-3, "converged": conv_all,
                   "note": "termination rule of harmonic_execute_gpu; excludes H2D/D2H; every tile swept in every pass"}
            sec, it, delta, conv, skipped = run_to_epsilon(True)
            # correctness evidence that travels with the number: the converged field, hashed band by band on the
            # rank that owns it, compared with the hash committed for this workload at N = 1
            digests = band_digests(slab.download_owned(), slab.row0)
            skipped_all = [int(skipped)]
            if world > 1:
                gathered = [None] * world
                dist.all_gather_object(gathered, (digests, int(skipped)))
                digests = [d for part, _ in gathered for d in part]
                skipped_all = [sk for _, sk in gathered]
            tte["field_sha1"] = field_hash(digests)
            tte["tiles_skipped_by_rank"] = skipped_all
            if it_all is None:
                it_all, delta_all = it, delta
            tte["with_static_tile_skipping"] = {
                "seconds": sec, "iterations": it, "delta": delta,
                "identical_to_all_tiles_run": bool(it == it_all and delta == delta_all),
                "tiles_skipped_rank0": int(skipped),
                "note": "the library's own loop (Field::solve with CUDA-graph periods at N = 1, ShardedSolver.solve at N > 1): "
                        "tiles whose 3x3 neighbourhood saw no update change a value in the previous pass return at once "
                        "(tiles reading ghost rows always run); results are bit-identical"}
        slab.field.close()
        return {"value": value, "ms_per_step": ms / args.steps, "math": math,
                "config_extra": {"math": math, "sweeps_per_pass": info["sweeps_per_pass"], "tile_rows": info["tile_rows"]},
                "clocks": clocks.summary(),
                "e2e": {"value": e2e_value, "unit": "Gcell-updates/s", "h2d_bytes_per_step": int(h2d) * world,
                        "d2h_bytes_per_step": int(d2h) * world, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                        "path": "libepic C ABI (update_model -> update_and_check + 99 x update -> "
                                "get_potential_values), pinned host arrays" if world == 1 else
                                "slab API per rank (upload -> 100 sweeps with halo exchange -> download), pinned"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved * world, "peak": peak * world, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "kernel": "sweep%dd_kernel<%s>" % (args.dims, "StrictMath" if math == "strict" else "FastMath"),
                             "kernel_ms": kern_ms, "algorithmic_bytes_per_update": ALGO_BYTES_PER_UPDATE,
                             "updates_per_launch": own_updates_per_pass},
                "time_to_epsilon": tte, "delta_after_timed_steps": delta_after}

    main_res = measure(args.math, args.tte)
    other = None
    if not args.single_mode:
        other = measure("fast" if args.math == "strict" else "strict", args.tte and args.tte_both_modes)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_rows = size if args.dims == 2 else min(size, 256)
        rate, dt, kind, cores, what = cpu_reference_rate(args, cpu_rows, 4)
        cpu = {"value": rate / 1e9, "unit": "Gcell-updates/s", "cores": cores, "kind": kind, "host_cores": os.cpu_count(),
               "sample": "4 half-sweeps of x0 = 0..%d of the %s grid (%.1f s); %s" % (
                   cpu_rows - 1, "x".join(str(v) for v in shape), dt, what)}

    if world > 1:
        barrier()
        dist.destroy_process_group()
    if rank != 0:
        return      # the other ranks are done: rank 0 goes on alone with every GPU of the node
    del u_pin, l_pin

    gpu_base = None
    if world == 1 and not args.no_gpu_baseline:
        gpu_base = reference_gpu_baseline(args)
    abi = None
    if world > 1 and not args.no_abi_multi and size * layer_cells <= 2 ** 30:
        try:
            abi = abi_multi(args, world, updates_per_step)
        except Exception as e:      # noqa: BLE001 -- the distributed numbers above stand on their own
            abi = {"error": str(e)[:300]}

    # The headline e2e is the call a user of the reference makes: the libepic C ABI.  At N = 1 the ranks' leg already
    # is that call; at N > 1 it is the single-process EPIC_DEVICES leg, and the per-rank slab API figure stays beside it.
    if abi is not None and "e2e" in abi:
        main_res["e2e_ranks"] = main_res["e2e"]
        main_res["e2e"] = dict(abi["e2e"], path="libepic C ABI, one process, EPIC_DEVICES naming all %d GPUs: " % world + abi["e2e"]["path"])
    tte = main_res["time_to_epsilon"] or {}
    skip = tte.get("with_static_tile_skipping") or {}
    key = "%s | %s" % (workload_name(args), main_res["math"])
    want = committed_fields().get(key)
    line = {"metric": "Gcell-updates/s", "value": main_res["value"], "unit": "Gcell-updates/s", "n_gpus": world,
            # compact keys first: what a scaling table needs besides `value`
            "tte_s": tte.get("seconds"), "tte_skip_s": skip.get("seconds"), "tte_iterations": skip.get("iterations"),
            "field_sha1": tte.get("field_sha1"),
            "field_sha1_matches_n1": (None if want is None or tte.get("field_sha1") is None else
                                      bool(want["field_sha1"] == tte["field_sha1"] and want["iterations"] == skip.get("iterations"))),
            "roofline_frac": main_res["roofline"]["frac"], "e2e_value": main_res["e2e"]["value"],
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload(args), **main_res["config_extra"]),
            "clocks": main_res["clocks"], "e2e": main_res["e2e"], "gpu_launches": main_res["gpu_launches"],
            "roofline": main_res["roofline"], "cpu_baseline": cpu, "gpu_baseline": gpu_base, "abi_multi": abi,
            "e2e_ranks": main_res.get("e2e_ranks"),
            "time_to_epsilon": main_res["time_to_epsilon"],
            "delta_after_timed_steps": main_res["delta_after_timed_steps"],
            "modes": "strict = bit-identical to the reference CPU path (default of the library; fields, deltas, iteration "
                     "counts and streamlines equal harmonic_complete_cpu's bit for bit); fast = TOLERANCE mode: MUFU "
                     "ex2/lg2, the arithmetic of the reference's own GPU kernel, |du| <= 1e-5*|u| + 4e-7*iterations at "
                     "matched epsilon, streamlines within 0.05 cell but NOT cell-for-cell",
            "library": lib.epic_b200_version().decode()}
    if other is not None:
        o_tte = other.get("time_to_epsilon") or {}
        o_skip = o_tte.get("with_static_tile_skipping") or {}
        line[other["math"] + "_mode"] = {k: other[k] for k in ("value", "ms_per_step", "e2e", "roofline", "clocks",
                                                               "gpu_launches")}
        line[other["math"] + "_mode"].update({"tte_skip_s": o_skip.get("seconds"), "tte_iterations": o_skip.get("iterations"),
                                              "field_sha1": o_tte.get("field_sha1")})
    print(json.dumps(line), flush=True)


def abi_multi(args, world, updates_per_step):
    """The same workload through the libepic C ABI alone, one process, EPIC_DEVICES naming all `world` GPUs: e2e
    steps with host buffers, then harmonic_execute_gpu to epsilon (the device seconds of its loop come from the
    library's statistics export; the ABI call itself also downloads the field)."""
    import ctypes as ct

    import torch

    from epic_b200 import libepic
    from epic_b200.harmonic import Harmonic
    os.environ["EPIC_DEVICES"] = ",".join(str(i) for i in range(world))
    os.environ["EPIC_MATH"] = args.math
    u, locked = make_grid(args)
    u_pin = torch.from_numpy(u).pin_memory()
    l_pin = torch.from_numpy(locked.view(np.int32)).pin_memory()
    del u, locked
    u_host, l_host = u_pin.numpy(), l_pin.numpy().view(np.uint32)
    u_keep = u_host.copy()
    h = Harmonic(u_host, l_host, 1e-3, SWEEPS_PER_STEP)
    h.initialize_gpu()
    out = {"path": "libepic C ABI, single process, EPIC_DEVICES=%s (engine/grid.cu)" % os.environ["EPIC_DEVICES"],
           "math": args.math}

    def e2e_step():
        h.currentIteration = 0
        h.update_model_gpu()
        h.run_iterations(SWEEPS_PER_STEP, "gpu")
        h.get_potential_values_gpu()
    e2e_step()
    steps = 3
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()      # uploads whatever the previous step downloaded: the bytes moved are the same
    dt = time.perf_counter() - t0
    out["e2e"] = {"value": updates_per_step * steps / dt / 1e9, "unit": "Gcell-updates/s", "steps": steps,
                  "ms_per_step": dt / steps * 1e3, "h2d_bytes_per_step": int(u_host.nbytes + l_host.nbytes),
                  "d2h_bytes_per_step": int(u_host.nbytes),
                  "path": "update_model -> update_and_check + 99 x update -> get_potential_values, pinned host arrays, "
                          "one host thread per slab for the copies"}
    # device-resident steps through the update / update_and_check calls
    h.currentIteration = 1
    h.run_iterations(SWEEPS_PER_STEP, "gpu")
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        h.run_iterations(SWEEPS_PER_STEP, "gpu")      # ends in a check sweep: synchronous
    dt = time.perf_counter() - t0
    out["value"] = updates_per_step * max(1, args.steps) / dt / 1e9
    # to epsilon
    u_host[:] = u_keep
    h.currentIteration = 0
    h.update_model_gpu()
    t0 = time.perf_counter()
    r = libepic.load().harmonic_execute_gpu(ct.byref(h), 1024)
    wall = time.perf_counter() - t0
    st = h.gpu_stats()
    digests = band_digests(u_host, 0)
    out["tte"] = {"return": int(r), "seconds": st["last_solve_seconds"], "execute_gpu_wall_seconds": wall,
                  "iterations": int(h.currentIteration), "delta": float(h.delta), "slabs": st["slabs"],
                  "tiles_skipped_by_slab": st["skipped_tiles"], "field_sha1": field_hash(digests)}
    h.uninitialize_gpu()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--size", type=int, default=16384, help="cells per dimension")
    ap.add_argument("--dims", type=int, choices=[2, 3], default=2,
                    help="2: BASELINE.json config 3 (size^2, the default and the headline); 3: config 5 (use --size 1024)")
    ap.add_argument("--math", choices=["strict", "fast"], default=os.environ.get("EPIC_MATH", "strict"))
    ap.add_argument("--tte", action="store_true", default=True,
                    help="also run to epsilon and report the time (default; twice: every tile swept / static tiles skipped)")
    ap.add_argument("--no-tte", dest="tte", action="store_false", help="skip the time-to-epsilon runs")
    ap.add_argument("--no-tte-all-tiles", dest="tte_all_tiles", action="store_false", default=True,
                    help="only the solve with static-tile skipping (halves the run time of the tte leg)")
    ap.add_argument("--tte-both-modes", action="store_true", help="time-to-epsilon for the other arithmetic mode as well")
    ap.add_argument("--workload", choices=["random", "maze"], default="random",
                    help="random: BASELINE.json config 3 / 5 (random obstacles); maze: config 4 (procedural maze, use --size 65536 --gpus 8)")
    ap.add_argument("--corridor", type=int, default=8, help="corridor width of the procedural maze")
    ap.add_argument("--tte-maze", action="store_true",
                    help="run the maze workload to epsilon as well (only sensible for small sizes; bounded by --tte-max-iterations)")
    ap.add_argument("--goals", type=int, default=None, help="goal cells (default 64 random-obstacle, 4 maze)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference-GPU leg (N = 1)")
    ap.add_argument("--try-stock-reference-gpu", action="store_true",
                    help="also attempt the reference's STOCK GPU build (it deadlocks on sm_100a; 25-second timeout)")
    ap.add_argument("--no-abi-multi", action="store_true", help="skip the single-process EPIC_DEVICES leg (N > 1)")
    ap.add_argument("--tte-max-iterations", type=int, default=400000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--halo", choices=["p2p", "nccl"], default="p2p", help="halo transport for --gpus > 1")
    ap.add_argument("--single-mode", action="store_true", help="measure only --math, not the other mode as well")
    args = ap.parse_args()
    if args.goals is None:
        args.goals = 64 if args.workload == "random" else 4
    if args.workload == "maze" and not args.tte_maze:
        # A relaxation needs O(L^2) iterations for a corridor path of L cells: the 482^2 demo maze takes 49 301, a
        # 65536^2 maze would take 10^9+.  Config 4 is a fixed-iteration throughput measurement unless asked otherwise.
        args.tte = False
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
