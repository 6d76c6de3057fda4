#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE ONLY -- times the UNTOUCHED reference GPU path recompiled for sm_100a
(oracle/_ref/libepic_ref_gpu.so, built by `make -C oracle refgpu` from the sources where they lie under
/root/reference) on this box's B200.  It is the same-box GPU baseline SURVEY.md section 2.2 names ("the bar
for the three 2D kernels is the reference source recompiled for sm_100a"); the product never loads it.

  python -m oracle.ref_gpu sweeps   --size 16384 --steps 20 --warmup 3     K x harmonic_update_gpu
  python -m oracle.ref_gpu complete --map maze                              harmonic_complete_gpu to epsilon

Each prints ONE JSON line.  bench.py runs this module in a subprocess under a timeout (the reference's
kernels call __syncthreads() inside divergent code, reference harmonic_gpu.cu:46-49; a hang must not take
the bench with it).  Note: the reference's GPU kernels use the OPPOSITE colour phase to its CPU path
(harmonic_gpu.cu:42-44 vs harmonic_cpu.cpp:49-51) and MUFU arithmetic, so iteration counts and fields are
those of the reference GPU path, not of the parity target; this is a timing baseline only.
"""
import argparse
import ctypes as ct
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REF_GPU_SO = os.path.join(HERE, "_ref", "libepic_ref_gpu.so")
# the same sources with __syncthreads() compiled out (oracle/Makefile, target refgpu_nobar): the stock build's
# divergent barriers deadlock on Volta and later, see profiles/r02_reference_gpu.md
REF_GPU_NOBAR_SO = os.path.join(HERE, "_ref", "libepic_ref_gpu_nobar.so")
_variant = "stock"


def available(variant="stock"):
    return os.path.exists(REF_GPU_SO if variant == "stock" else REF_GPU_NOBAR_SO)


def _lib():
    from oracle.oracle import RefHarmonic
    R = ct.CDLL(REF_GPU_SO if _variant == "stock" else REF_GPU_NOBAR_SO)
    P = ct.POINTER(RefHarmonic)
    for name in ("harmonic_initialize_dimension_size_gpu", "harmonic_initialize_potential_values_gpu",
                 "harmonic_initialize_locked_gpu", "harmonic_uninitialize_dimension_size_gpu",
                 "harmonic_uninitialize_potential_values_gpu", "harmonic_uninitialize_locked_gpu",
                 "harmonic_get_potential_values_gpu", "harmonic_uninitialize_gpu"):
        getattr(R, name).argtypes = (P,)
    for name in ("harmonic_complete_gpu", "harmonic_initialize_gpu", "harmonic_update_gpu",
                 "harmonic_update_and_check_gpu", "harmonic_execute_gpu"):
        getattr(R, name).argtypes = (P, ct.c_uint)
    return R


def _harmonic(u, locked, eps, stagger):
    from oracle.oracle import RefHarmonic
    h = RefHarmonic()
    m = np.array(u.shape, dtype=np.uint32)
    h.n = u.ndim
    h.m = m.ctypes.data_as(ct.POINTER(ct.c_uint))
    h.u = u.ctypes.data_as(ct.POINTER(ct.c_float))
    h.locked = locked.ctypes.data_as(ct.POINTER(ct.c_uint))
    h.epsilon, h.delta = eps, 0.0
    h.numIterationsToStaggerCheck, h.currentIteration = stagger, 0
    return h, m


def sweeps(size, steps, warmup, threads=1024):
    """GCUPS of `steps` harmonic_update_gpu calls (each: one kernel + cudaDeviceSynchronize, reference
    harmonic_gpu.cu:327-355) on the bench's random-obstacle grid, device-resident."""
    from epic_b200 import grids
    u, locked = grids.random_obstacles((size, size), 0.2, 64, seed=1234)
    R = _lib()
    h, _m = _harmonic(u, locked, 1e-3, 100)
    r = R.harmonic_initialize_dimension_size_gpu(ct.byref(h))
    r += R.harmonic_initialize_potential_values_gpu(ct.byref(h))
    r += R.harmonic_initialize_locked_gpu(ct.byref(h))
    r += R.harmonic_initialize_gpu(ct.byref(h), threads)
    if r != 0:
        return {"error": "initialize returned %d" % r}
    if _variant == "stock":
        R.harmonic_update_and_check_gpu(ct.byref(h), threads)
    for _ in range(warmup):
        R.harmonic_update_gpu(ct.byref(h), threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        r += R.harmonic_update_gpu(ct.byref(h), threads)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    rc = R.harmonic_update_and_check_gpu(ct.byref(h), threads) if _variant == "stock" else -1
    dt_check = time.perf_counter() - t1
    R.harmonic_uninitialize_gpu(ct.byref(h))
    R.harmonic_uninitialize_dimension_size_gpu(ct.byref(h))
    R.harmonic_uninitialize_potential_values_gpu(ct.byref(h))
    R.harmonic_uninitialize_locked_gpu(ct.byref(h))
    updates = float(size) * size / 2.0 * steps
    return {"kind": "reference GPU kernels (harmonic_gpu.cu) recompiled for sm_100a, stock host loop" +
                    ("" if _variant == "stock" else "; __syncthreads() compiled out (the stock build deadlocks)"),
            "what": "%d x harmonic_update_gpu (1 half-sweep + cudaDeviceSynchronize each), %dx%d random-obstacle grid, "
                    "numThreads %d" % (steps, size, size, threads),
            "value": updates / dt / 1e9, "unit": "Gcell-updates/s", "ms_per_half_sweep": dt / steps * 1e3,
            "ms_per_check_sweep": dt_check * 1e3, "errors": int(r), "check_return": int(rc)}


def complete(name, threads=1024):
    """Wall time of harmonic_complete_gpu (H2D + solve to epsilon + D2H, as the reference's API does it)."""
    from epic_b200 import grids
    maps = np.load(os.path.join(ROOT, "tests", "golden", "maps.npz"))
    u, locked = grids.grid_from_image(maps[name])
    R = _lib()
    h, _m = _harmonic(u, locked, 1e-3, 100)
    # one throw-away call on a copy: CUDA context creation and module load are not part of the solve
    u2, l2 = u.copy(), locked.copy()
    h2, _m2 = _harmonic(u2, l2, 1e-1, 100)
    R.harmonic_complete_gpu(ct.byref(h2), threads)
    t0 = time.perf_counter()
    r = R.harmonic_complete_gpu(ct.byref(h), threads)
    dt = time.perf_counter() - t0
    return {"kind": "reference GPU path recompiled for sm_100a", "what": "harmonic_complete_gpu on maps '%s' %s, eps 1e-3"
            % (name, "x".join(str(s) for s in u.shape)), "seconds": dt, "iterations": int(h.currentIteration),
            "delta": float(h.delta), "return": int(r),
            "gcups": float(u.size) / 2.0 * h.currentIteration / dt / 1e9}


def updates(name, iterations, threads=1024):
    """Wall time of `iterations` harmonic_update_gpu calls on a demo map, device-resident (H2D before, D2H after,
    both timed separately): what the reference's host loop + kernel cost for that many half-sweeps.  No
    convergence checks (the nobar build's check kernel is unusable), so this is a LOWER bound of a solve."""
    from epic_b200 import grids
    maps = np.load(os.path.join(ROOT, "tests", "golden", "maps.npz"))
    u, locked = grids.grid_from_image(maps[name])
    R = _lib()
    h, _m = _harmonic(u, locked, 1e-3, 100)
    t0 = time.perf_counter()
    r = R.harmonic_initialize_dimension_size_gpu(ct.byref(h))
    r += R.harmonic_initialize_potential_values_gpu(ct.byref(h))
    r += R.harmonic_initialize_locked_gpu(ct.byref(h))
    t_up = time.perf_counter() - t0
    if r != 0:
        return {"error": "initialize returned %d" % r}
    for _ in range(200):
        R.harmonic_update_gpu(ct.byref(h), threads)
    t0 = time.perf_counter()
    for _ in range(iterations):
        r += R.harmonic_update_gpu(ct.byref(h), threads)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    R.harmonic_get_potential_values_gpu(ct.byref(h))
    t_down = time.perf_counter() - t0
    R.harmonic_uninitialize_dimension_size_gpu(ct.byref(h))
    R.harmonic_uninitialize_potential_values_gpu(ct.byref(h))
    R.harmonic_uninitialize_locked_gpu(ct.byref(h))
    return {"kind": "reference GPU kernels recompiled for sm_100a" +
                    ("" if _variant == "stock" else "; __syncthreads() compiled out (the stock build deadlocks)"),
            "what": "%d x harmonic_update_gpu on maps '%s' %s (the iteration count of the reference CPU solve at eps 1e-3); "
                    "no convergence checks" % (iterations, name, "x".join(str(s) for s in u.shape)),
            "seconds": dt, "us_per_half_sweep": dt / iterations * 1e6, "upload_seconds": t_up, "download_seconds": t_down,
            "errors": int(r), "gcups": float(u.size) / 2.0 * iterations / dt / 1e9}


def main():
    global _variant
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["sweeps", "complete", "updates"])
    ap.add_argument("--variant", choices=["stock", "nobar"], default="stock")
    ap.add_argument("--iterations", type=int, default=49301)
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--map", default="maze")
    ap.add_argument("--threads", type=int, default=1024)
    a = ap.parse_args()
    _variant = a.variant
    if not available(a.variant):
        print(json.dumps({"unavailable": "oracle/_ref/libepic_ref_gpu*.so was not built (no /root/reference at build time)"}))
        return
    if a.mode == "sweeps":
        out = sweeps(a.size, a.steps, a.warmup, a.threads)
    elif a.mode == "complete":
        out = complete(a.map, a.threads)
    else:
        out = updates(a.map, a.iterations, a.threads)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
