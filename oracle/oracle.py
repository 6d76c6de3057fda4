"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for oracle/_build/liboracle.so (the C
restatement, oracle/harmonic_oracle.c) and, when it was built in the container,
oracle/_ref/libepic_ref_cpu.so (the untouched reference CPU sources).

Nothing here is imported by the product package.  See harmonic_oracle.h for how the
restatement is pinned against the reference.
"""
import ctypes as ct
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libepic_ref_cpu.so")

SUCCESS, CONVERGED = 0, 1
INVALID_DATA, INVALID_LOCATION, INVALID_GRADIENT, INVALID_PATH = 2, 10, 12, 13


def build(force=False):
    """Compile the restatement (and the reference CPU sources when /root/reference exists)."""
    if force or not os.path.exists(ORACLE_SO):
        subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/libepic/src/harmonic") and (force or not os.path.exists(REF_SO)):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


class _OH(ct.Structure):
    _fields_ = [("n", ct.c_uint32), ("m", ct.c_uint64 * 4), ("u", ct.POINTER(ct.c_float)),
                ("locked", ct.POINTER(ct.c_uint32)), ("epsilon", ct.c_float), ("delta", ct.c_float),
                ("stagger", ct.c_uint32), ("iteration", ct.c_uint32), ("threads", ct.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ct.CDLL(ORACLE_SO)
        P = ct.POINTER(_OH)
        L.oracle_sweep.argtypes = (P, ct.c_int)
        L.oracle_sweep.restype = None
        for name in ("oracle_update", "oracle_update_and_check", "oracle_complete"):
            getattr(L, name).argtypes = (P,)
        L.oracle_run_iterations.argtypes = (P, ct.c_uint32)
        L.oracle_set_cells_2d.argtypes = (P, ct.c_uint32, ct.POINTER(ct.c_uint32), ct.POINTER(ct.c_uint32))
        L.oracle_potential_2d.argtypes = (P, ct.c_float, ct.c_float, ct.POINTER(ct.c_float))
        L.oracle_gradient_2d.argtypes = (P, ct.c_float, ct.c_float, ct.c_float,
                                         ct.POINTER(ct.c_float), ct.POINTER(ct.c_float))
        L.oracle_path_2d.argtypes = (P, ct.c_float, ct.c_float, ct.c_float, ct.c_float, ct.c_uint32,
                                     ct.POINTER(ct.c_uint32), ct.POINTER(ct.POINTER(ct.c_float)))
        L.oracle_free_path.argtypes = (ct.POINTER(ct.c_float),)
        L.oracle_free_path.restype = None
        _lib = L
    return _lib


class Oracle:
    """A grid held by the C restatement.  `u` (float32) and `locked` (uint32) are numpy
    arrays of shape `m`, modified in place."""

    def __init__(self, u, locked, epsilon=1e-3, stagger=100, threads=1):
        self.u = np.ascontiguousarray(u, dtype=np.float32)
        self.locked = np.ascontiguousarray(locked, dtype=np.uint32)
        assert self.u.shape == self.locked.shape and self.u.ndim in (2, 3)
        self.h = _OH()
        self.h.n = self.u.ndim
        for i, s in enumerate(self.u.shape):
            self.h.m[i] = s
        self.h.u = self.u.ctypes.data_as(ct.POINTER(ct.c_float))
        self.h.locked = self.locked.ctypes.data_as(ct.POINTER(ct.c_uint32))
        self.h.epsilon = epsilon
        self.h.delta = 0.0
        self.h.stagger = stagger
        self.h.iteration = 0
        self.h.threads = threads

    iteration = property(lambda s: s.h.iteration, lambda s, v: setattr(s.h, "iteration", v))
    delta = property(lambda s: s.h.delta)

    def update(self):
        return lib().oracle_update(ct.byref(self.h))

    def update_and_check(self):
        return lib().oracle_update_and_check(ct.byref(self.h))

    def complete(self):
        return lib().oracle_complete(ct.byref(self.h))

    def run_iterations(self, count):
        return lib().oracle_run_iterations(ct.byref(self.h), count)

    def set_cells(self, v, types):
        v = np.ascontiguousarray(v, dtype=np.uint32).reshape(-1)
        types = np.ascontiguousarray(types, dtype=np.uint32)
        return lib().oracle_set_cells_2d(ct.byref(self.h), len(types),
                                         v.ctypes.data_as(ct.POINTER(ct.c_uint32)),
                                         types.ctypes.data_as(ct.POINTER(ct.c_uint32)))

    def potential(self, x, y):
        out = ct.c_float(0.0)
        r = lib().oracle_potential_2d(ct.byref(self.h), x, y, ct.byref(out))
        return r, out.value

    def gradient(self, x, y, cd):
        px, py = ct.c_float(0.0), ct.c_float(0.0)
        r = lib().oracle_gradient_2d(ct.byref(self.h), x, y, cd, ct.byref(px), ct.byref(py))
        return r, px.value, py.value

    def path(self, x, y, step, cd, max_length):
        k = ct.c_uint32(0)
        p = ct.POINTER(ct.c_float)()
        r = lib().oracle_path_2d(ct.byref(self.h), x, y, step, cd, max_length, ct.byref(k), ct.byref(p))
        if r != SUCCESS:
            return r, np.zeros((0, 2), np.float32)
        out = np.ctypeslib.as_array(p, shape=(2 * k.value,)).copy().reshape(-1, 2)
        lib().oracle_free_path(p)
        return r, out


# ---------------------------------------------------------------------------------------
# The untouched reference (container only).  Struct layout: reference
# libepic/include/epic/harmonic/harmonic.h:44-64 (80 bytes on x86-64).

class RefHarmonic(ct.Structure):
    _fields_ = [("n", ct.c_uint), ("m", ct.POINTER(ct.c_uint)), ("u", ct.POINTER(ct.c_float)),
                ("locked", ct.POINTER(ct.c_uint)), ("epsilon", ct.c_float), ("delta", ct.c_float),
                ("numIterationsToStaggerCheck", ct.c_uint), ("currentIteration", ct.c_uint),
                ("d_m", ct.POINTER(ct.c_uint)), ("d_u", ct.POINTER(ct.c_float)),
                ("d_locked", ct.POINTER(ct.c_uint)), ("d_delta", ct.POINTER(ct.c_float))]


def have_ref():
    build()
    return os.path.exists(REF_SO)


_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        build()
        R = ct.CDLL(REF_SO)
        P = ct.POINTER(RefHarmonic)
        for name in ("harmonic_complete_cpu", "harmonic_update_cpu", "harmonic_update_and_check_cpu"):
            getattr(R, name).argtypes = (P,)
        R.harmonic_utilities_set_cells_2d_cpu.argtypes = (P, ct.c_uint, ct.POINTER(ct.c_uint), ct.POINTER(ct.c_uint))
        R.harmonic_compute_potential_2d_cpu.argtypes = (P, ct.c_float, ct.c_float, ct.POINTER(ct.c_float))
        R.harmonic_compute_gradient_2d_cpu.argtypes = (P, ct.c_float, ct.c_float, ct.c_float,
                                                       ct.POINTER(ct.c_float), ct.POINTER(ct.c_float))
        R.harmonic_compute_path_2d_cpu.argtypes = (P, ct.c_float, ct.c_float, ct.c_float, ct.c_float, ct.c_uint,
                                                   ct.POINTER(ct.c_uint), ct.POINTER(ct.POINTER(ct.c_float)))
        R.harmonic_free_path_cpu.argtypes = (ct.POINTER(ct.POINTER(ct.c_float)),)
        _ref = R
    return _ref


class Reference:
    """The same interface as `Oracle`, served by the compiled reference sources."""

    def __init__(self, u, locked, epsilon=1e-3, stagger=100):
        self.u = np.ascontiguousarray(u, dtype=np.float32)
        self.locked = np.ascontiguousarray(locked, dtype=np.uint32)
        self.m = np.array(self.u.shape, dtype=np.uint32)
        self.h = RefHarmonic()
        self.h.n = self.u.ndim
        self.h.m = self.m.ctypes.data_as(ct.POINTER(ct.c_uint))
        self.h.u = self.u.ctypes.data_as(ct.POINTER(ct.c_float))
        self.h.locked = self.locked.ctypes.data_as(ct.POINTER(ct.c_uint))
        self.h.epsilon = epsilon
        self.h.delta = 0.0
        self.h.numIterationsToStaggerCheck = stagger
        self.h.currentIteration = 0

    iteration = property(lambda s: s.h.currentIteration, lambda s, v: setattr(s.h, "currentIteration", v))
    delta = property(lambda s: s.h.delta)

    def update(self):
        return ref_lib().harmonic_update_cpu(ct.byref(self.h))

    def update_and_check(self):
        return ref_lib().harmonic_update_and_check_cpu(ct.byref(self.h))

    def complete(self):
        return ref_lib().harmonic_complete_cpu(ct.byref(self.h))

    def run_iterations(self, count):
        for _ in range(count):
            if self.h.currentIteration % self.h.numIterationsToStaggerCheck == 0:
                self.update_and_check()
            else:
                self.update()
        return SUCCESS

    def set_cells(self, v, types):
        v = np.ascontiguousarray(v, dtype=np.uint32).reshape(-1)
        types = np.ascontiguousarray(types, dtype=np.uint32)
        return ref_lib().harmonic_utilities_set_cells_2d_cpu(
            ct.byref(self.h), len(types), v.ctypes.data_as(ct.POINTER(ct.c_uint)),
            types.ctypes.data_as(ct.POINTER(ct.c_uint)))

    def potential(self, x, y):
        out = ct.c_float(0.0)
        r = ref_lib().harmonic_compute_potential_2d_cpu(ct.byref(self.h), x, y, ct.byref(out))
        return r, out.value

    def gradient(self, x, y, cd):
        px, py = ct.c_float(0.0), ct.c_float(0.0)
        r = ref_lib().harmonic_compute_gradient_2d_cpu(ct.byref(self.h), x, y, cd, ct.byref(px), ct.byref(py))
        return r, px.value, py.value

    def path(self, x, y, step, cd, max_length):
        k = ct.c_uint(0)
        p = ct.POINTER(ct.c_float)()
        r = ref_lib().harmonic_compute_path_2d_cpu(ct.byref(self.h), x, y, step, cd, max_length,
                                                   ct.byref(k), ct.byref(p))
        if r != SUCCESS:
            return r, np.zeros((0, 2), np.float32)
        out = np.ctypeslib.as_array(p, shape=(2 * k.value,)).copy().reshape(-1, 2)
        ref_lib().harmonic_free_path_cpu(ct.byref(p))
        return r, out
