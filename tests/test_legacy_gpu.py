"""the reference's legacy linear-space SOR (harmonic_legacy_cpu.cpp:36-141, host only in the
reference) on the GPU.  The lexicographic in-place Gauss-Seidel sweep is executed as waves of independent cells
(epic_b200/csrc/abi/legacy_gpu.cu); iteration count and every bit of the field must equal the library's CPU export,
which tests/test_abi.py and the golden replay pin to the reference's own code."""
import ctypes as ct

import numpy as np
import pytest

from epic_b200 import grids

pytestmark = pytest.mark.gpu


def run(lib, name, which, w, h, eps, omega, locked, u):
    ctype = ct.c_float if name == "float" else ct.c_double
    it = ct.c_uint(0)
    fn = getattr(lib, "harmonic_legacy_sor_2d_%s_%s" % (name, which))
    r = fn(w, h, ctype(eps), ctype(omega), locked.ctypes.data_as(ct.POINTER(ct.c_uint)),
           u.ctypes.data_as(ct.POINTER(ctype)), ct.byref(it))
    return r, it.value


@pytest.mark.parametrize("name,dtype,eps", [("float", np.float32, 1e-4), ("double", np.float64, 1e-9)])
@pytest.mark.parametrize("shape,omega,seed", [((37, 53), 1.5, 1), ((96, 130), 1.5, 2), ((64, 64), 1.0, 3), ((5, 200), 1.8, 4)])
def test_legacy_sor_gpu_equals_cpu_export(libepic_built, name, dtype, eps, shape, omega, seed):
    """Linear-space convention of the legacy solver: obstacles locked at 1, goals locked at 0, free cells start at 1."""
    h, w = shape
    _, locked = grids.random_obstacles(shape, 0.15, 3, seed=seed)
    rng = np.random.RandomState(seed)
    u0 = np.ones(shape, dtype)
    ys, xs = np.nonzero(locked[1:-1, 1:-1] == 0)
    for i in rng.choice(len(ys), 3, replace=False):
        u0[ys[i] + 1, xs[i] + 1] = 0.0
        locked[ys[i] + 1, xs[i] + 1] = 1
    free = locked == 0
    u0[free] = rng.random_sample(int(free.sum())).astype(dtype)     # arbitrary start: every cell moves at once
    uc, ug = u0.copy(), u0.copy()
    rc, itc = run(libepic_built, name, "cpu", w, h, eps, omega, locked, uc)
    rg, itg = run(libepic_built, name, "gpu", w, h, eps, omega, locked, ug)
    assert rc == 0 and rg == 0
    assert itg == itc and itc >= 10000
    assert np.array_equal(ug, uc), "max |diff| %g" % np.abs(ug.astype(np.float64) - uc.astype(np.float64)).max()
    assert not np.array_equal(ug, u0)


def test_legacy_sor_gpu_beyond_the_minimum_iteration_count(libepic_built):
    """A tight epsilon on an open 120 x 160 room (plain Gauss-Seidel, omega = 1) makes the loop run twice its
    10000-iteration minimum: the stopping iteration lies in the sixth window of the discovery run."""
    h, w = 120, 160
    locked = np.ones((h, w), np.uint32)
    locked[1:-1, 1:-1] = 0
    u0 = np.ones((h, w), np.float64)
    u0[h // 2, w // 2] = 0.0
    locked[h // 2, w // 2] = 1
    uc, ug = u0.copy(), u0.copy()
    rc, itc = run(libepic_built, "double", "cpu", w, h, 1e-10, 1.0, locked, uc)
    rg, itg = run(libepic_built, "double", "gpu", w, h, 1e-10, 1.0, locked, ug)
    assert rc == 0 and rg == 0 and itc > 20000
    assert itg == itc and np.array_equal(ug, uc)


def test_legacy_sor_gpu_degenerate_and_invalid(libepic_built):
    locked = np.ones((2, 7), np.uint32)
    u = np.ones((2, 7), np.float32)
    r, it = run(libepic_built, "float", "gpu", 7, 2, 1e-3, 1.5, locked, u)
    assert (r, it) == (0, 10000) and np.all(u == 1.0)
    it = ct.c_uint(0)
    assert libepic_built.harmonic_legacy_sor_2d_float_gpu(7, 2, ct.c_float(1e-3), ct.c_float(1.5), None,
                                                          u.ctypes.data_as(ct.POINTER(ct.c_float)), ct.byref(it)) == 2
