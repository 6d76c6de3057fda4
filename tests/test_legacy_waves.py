"""The claim behind epic_b200/csrc/abi/legacy_gpu.cu, checked on the CPU: the reference's lexicographic in-place SOR
sweep (harmonic_legacy_cpu.cpp:36-70) equals, bit for bit, the wave schedule the GPU kernel runs -- wave
w = (x - 1) + (y - 1) + 2k holds independent cells (every second anti-diagonal, each with its own iteration number k),
and updating the array in place wave after wave is the same computation.  numpy stands in for the kernel here; the
-m gpu tier compares the kernel itself with the library's CPU export (tests/test_legacy_gpu.py)."""
import numpy as np
import pytest


def lexicographic(u, locked, omega, iterations):
    u = u.copy()
    h, w = u.shape
    one_minus, quarter = u.dtype.type(1.0) - omega, omega / u.dtype.type(4.0)
    deltas = []
    for _ in range(iterations):
        delta = u.dtype.type(0.0)
        for y in range(1, h - 1):
            for x in range(1, w - 1):
                if locked[y, x] == 1:
                    continue
                before = u[y, x]
                s = u[y - 1, x] + u[y + 1, x]
                s = s + u[y, x - 1]
                s = s + u[y, x + 1]
                u[y, x] = one_minus * before + quarter * s
                delta = max(delta, abs(u[y, x] - before))
        deltas.append(delta)
    return u, deltas


def waves(u, locked, omega, iterations):
    """The kernel's schedule: all cells of a wave computed from the array as it stands, then written together."""
    u = u.copy()
    h, w = u.shape
    one_minus, quarter = u.dtype.type(1.0) - omega, omega / u.dtype.type(4.0)
    yi, xi = np.meshgrid(np.arange(h - 2), np.arange(w - 2), indexing="ij")
    d = yi + xi
    free = locked[1:-1, 1:-1] != 1
    deltas = np.zeros(iterations, u.dtype)
    for wave in range(int(d.max()) + 2 * (iterations - 1) + 1):
        k2 = wave - d
        active = free & (k2 >= 0) & (k2 % 2 == 0) & (k2 // 2 < iterations)
        if not active.any():
            continue
        inner = u[1:-1, 1:-1]
        s = u[:-2, 1:-1] + u[2:, 1:-1]
        s = s + u[1:-1, :-2]
        s = s + u[1:-1, 2:]
        new = one_minus * inner + quarter * s
        change = np.abs(new - inner)
        ks = (k2 // 2)[active]
        np.maximum.at(deltas, ks, change[active])
        inner[active] = new[active]
    return u, list(deltas)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,omega", [((9, 13), 1.5), ((12, 7), 1.0), ((5, 5), 1.8), ((3, 20), 1.3)])
def test_wave_schedule_equals_lexicographic_sweep(dtype, shape, omega):
    rng = np.random.RandomState(shape[0] * 31 + shape[1])
    locked = (rng.random_sample(shape) < 0.2).astype(np.uint32)
    locked[0, :] = locked[-1, :] = locked[:, 0] = locked[:, -1] = 1
    u0 = rng.random_sample(shape).astype(dtype)
    om = dtype(omega)
    a, da = lexicographic(u0, locked, om, 7)
    b, db = waves(u0, locked, om, 7)
    assert np.array_equal(a, b)
    assert [float(x) for x in da] == [float(x) for x in db]
