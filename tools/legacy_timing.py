#!/usr/bin/env python
"""Legacy linear-space SOR (double, omega 1.5, epsilon 1e-10 -- the setting of the reference's batch harness,
libepic/tests/batch/batch.py:52-70) on maps/maze.png: harmonic_legacy_sor_2d_double_cpu vs ..._gpu (extension)."""
import ctypes as ct
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402

import common  # noqa: E402
from epic_b200 import libepic as le  # noqa: E402

L = le.load()
for name in ("basic", "maze"):
    img = common.maps()[name]
    h, w = img.shape
    locked = ((img == 0) | (img == 255)).astype(np.uint32)
    u0 = np.where(img == 255, 0.0, 1.0).astype(np.float64)      # linear space: goal 0, everything else 1
    res = {}
    for which in ("gpu", "gpu", "cpu"):
        u = u0.copy()
        it = ct.c_uint(0)
        t0 = time.perf_counter()
        r = getattr(L, "harmonic_legacy_sor_2d_double_" + which)(w, h, ct.c_double(1e-10), ct.c_double(1.5),
                                                                 locked.ctypes.data_as(ct.POINTER(ct.c_uint)),
                                                                 u.ctypes.data_as(ct.POINTER(ct.c_double)), ct.byref(it))
        res[which] = (r, it.value, time.perf_counter() - t0, u)
    print("%-6s %s: iterations cpu %d gpu %d, cpu %.2f s, gpu %.3f s (discovery + solve), identical %s" % (
        name, img.shape, res["cpu"][1], res["gpu"][1], res["cpu"][2], res["gpu"][2], np.array_equal(res["cpu"][3], res["gpu"][3])), flush=True)
