#!/usr/bin/env python
"""Device-resident sweep throughput for a few configurations (CUDA events around N passes).
usage: sweep_timing.py size math [warm_sweeps] [passes]   (EPIC_THREADS / EPIC_TILE_ROWS / EPIC_SWEEPS_PER_PASS from env)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from epic_b200 import grids  # noqa: E402
from epic_b200.field import Field  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
math = sys.argv[2] if len(sys.argv) > 2 else "strict"
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 400
passes = int(sys.argv[4]) if len(sys.argv) > 4 else 40
_cache = {}


def run(threads, th, T=None):
    for k, v in (("EPIC_THREADS", threads), ("EPIC_TILE_ROWS", th), ("EPIC_SWEEPS_PER_PASS", T)):
        if v:
            os.environ[k] = str(v)
        else:
            os.environ.pop(k, None)
    if "grid" not in _cache:
        _cache["grid"] = grids.random_obstacles((size, size), 0.2, 64, seed=1234)
    u, locked = _cache["grid"]
    stream = torch.cuda.current_stream().cuda_stream
    f = Field((size, size), math=math, stream=stream)
    f.upload(u, locked)
    T_ = f.info()["sweeps_per_pass"]
    f.run(0, warm, False)
    f.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    f.run(warm, passes * T_, False)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    info = f.info()
    gcups = size * size / 2 * passes * T_ / (ms * 1e-3) / 1e9
    print("size %d math %-6s threads %s tile_rows %d T %d : %.3f ms/pass  %.1f GCUPS" % (
        size, math, threads or "auto", info["tile_rows"], T_, ms / passes, gcups), flush=True)
    f.close()


for cfg in (os.environ.get("CONFIGS") or "256:96,512:96,512:64,256:48").split(","):
    parts = cfg.split(":")
    run(int(parts[0]), int(parts[1]), int(parts[2]) if len(parts) > 2 else None)
