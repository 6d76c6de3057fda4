"""Python mirror of the reference's `Harmonic` wrapper (libepic/python/epic/harmonic.py:34-102) and of
the map loader / streamline part of `HarmonicMap` (harmonic_map.py:62-127), bound to this repository's
libepic.so.

Differences from the reference wrapper, on purpose: `solve(process='gpu')` does not silently fall
back to the CPU -- a GPU failure raises; the CPU solver runs only when the caller asks for
process='cpu' (it is the library's own `harmonic_complete_cpu` export).  Timing uses
time.perf_counter (the reference's time.clock no longer exists).
"""
import ctypes as ct
import time

import numpy as np

from . import grids
from . import libepic as le


class EpicError(RuntimeError):
    def __init__(self, function, code):
        RuntimeError.__init__(self, "%s returned %d" % (function, code))
        self.function, self.code = function, code


class Harmonic(le.EpicHarmonic):
    """A harmonic function over an n-dimensional occupancy grid, held by numpy arrays."""

    def __init__(self, u=None, locked=None, epsilon=1e-2, stagger=100):
        le.EpicHarmonic.__init__(self)
        self.n = 0
        self.epsilon = epsilon
        self.delta = epsilon + 1.0
        self.numIterationsToStaggerCheck = int(stagger)
        self.currentIteration = 0
        self._m = self._u = self._locked = None
        if u is not None:
            self.set_grid(u, locked)

    def set_grid(self, u, locked):
        self._u = np.ascontiguousarray(u, dtype=np.float32)
        self._locked = np.ascontiguousarray(locked, dtype=np.uint32)
        assert self._u.shape == self._locked.shape
        self._m = np.array(self._u.shape, dtype=np.uint32)
        self.n = self._u.ndim
        self.m = self._m.ctypes.data_as(ct.POINTER(ct.c_uint))
        self.u = self._u.ctypes.data_as(ct.POINTER(ct.c_float))
        self.locked = self._locked.ctypes.data_as(ct.POINTER(ct.c_uint))

    # numpy views of the host arrays the struct points at
    @property
    def field(self):
        return self._u

    @property
    def locked_cells(self):
        return self._locked

    def _call(self, name, *args, ok=(le.EPIC_SUCCESS,)):
        r = getattr(le.load(), name)(ct.byref(self), *args)
        if r not in ok:
            raise EpicError(name, r)
        return r

    # -- residency (reference harmonic_model_gpu.h) ------------------------------------------------
    def initialize_gpu(self):
        self._call("harmonic_initialize_dimension_size_gpu")
        self._call("harmonic_initialize_potential_values_gpu")
        self._call("harmonic_initialize_locked_gpu")

    def uninitialize_gpu(self):
        self._call("harmonic_uninitialize_dimension_size_gpu")
        self._call("harmonic_uninitialize_potential_values_gpu")
        self._call("harmonic_uninitialize_locked_gpu")

    def update_model_gpu(self):
        self._call("harmonic_update_model_gpu")

    def get_potential_values_gpu(self):
        self._call("harmonic_get_potential_values_gpu")
        return self._u

    def gpu_stats(self):
        """Extension: how the library holds this grid on the device(s) and what its last solve did (slabs,
        launches, device seconds / iterations / delta of the last harmonic_execute_gpu, tiles skipped per slab)."""
        st = le.GridStats()
        r = le.load().epic_b200_harmonic_stats(ct.cast(ct.byref(self), ct.c_void_p), ct.byref(st))
        if r != le.EPIC_SUCCESS:
            raise EpicError("epic_b200_harmonic_stats", r)
        return {"slabs": st.slabs, "launches": st.launches, "last_solve_seconds": st.last_solve_seconds,
                "last_solve_iterations": st.last_solve_iterations, "last_solve_delta": st.last_solve_delta,
                "skipped_tiles": [int(st.skipped_tiles[i]) for i in range(st.slabs)]}

    # -- solving -------------------------------------------------------------------------------------
    def solve(self, algorithm='gauss-seidel', process='gpu', numThreads=1024, epsilon=None):
        """Reference signature (harmonic.py:56).  Returns (wall seconds, cpu seconds) of the solver call."""
        if algorithm != 'gauss-seidel':
            raise ValueError("the algorithm '%s' is undefined" % algorithm)
        if epsilon is not None:
            self.epsilon = epsilon
        t0 = (time.perf_counter(), time.process_time())
        if process == 'gpu':
            self._call("harmonic_complete_gpu", int(numThreads))
        elif process == 'cpu':
            self._call("harmonic_complete_cpu")
        else:
            raise ValueError("process must be 'gpu' or 'cpu'")
        return time.perf_counter() - t0[0], time.process_time() - t0[1]

    def update_gpu(self, numThreads=1024):
        return self._call("harmonic_update_gpu", int(numThreads))

    def update_and_check_gpu(self, numThreads=1024):
        return self._call("harmonic_update_and_check_gpu", int(numThreads),
                          ok=(le.EPIC_SUCCESS, le.EPIC_SUCCESS_AND_CONVERGED))

    def update_cpu(self):
        return self._call("harmonic_update_cpu")

    def update_and_check_cpu(self):
        return self._call("harmonic_update_and_check_cpu", ok=(le.EPIC_SUCCESS, le.EPIC_SUCCESS_AND_CONVERGED))

    def run_iterations(self, count, process='gpu', numThreads=1024):
        """`count` iterations with the complete() schedule (check sweeps when currentIteration is a
        multiple of the stagger), without the termination test."""
        for _ in range(count):
            check = self.currentIteration % self.numIterationsToStaggerCheck == 0
            if process == 'gpu':
                self.update_and_check_gpu(numThreads) if check else self.update_gpu(numThreads)
            else:
                self.update_and_check_cpu() if check else self.update_cpu()

    def set_cells(self, v, types, process='gpu', numThreads=1024):
        v = np.ascontiguousarray(v, dtype=np.uint32).reshape(-1)
        types = np.ascontiguousarray(types, dtype=np.uint32)
        pv, pt = v.ctypes.data_as(ct.POINTER(ct.c_uint)), types.ctypes.data_as(ct.POINTER(ct.c_uint))
        if process == 'gpu':
            return self._call("harmonic_utilities_set_cells_2d_gpu", int(numThreads), len(types), pv, pt)
        return self._call("harmonic_utilities_set_cells_2d_cpu", len(types), pv, pt)

    def set_occupancy_grid(self, data, process='gpu', obstacleThreshold=50, noChangeValue=-2):
        """Dense map ingest (extension): `data` is a nav_msgs/OccupancyGrid-style int8 array of the grid's shape."""
        data = np.ascontiguousarray(data, dtype=np.int8)
        assert data.shape == self.field.shape
        return self._call("harmonic_utilities_set_occupancy_grid_2d_" + process, data.ctypes.data_as(ct.POINTER(ct.c_byte)),
                          int(obstacleThreshold), int(noChangeValue))

    def reset_free_cells(self, process='gpu'):
        return self._call("harmonic_utilities_reset_free_cells_2d_" + process)

    # -- streamlines ---------------------------------------------------------------------------------
    def compute_potential(self, x, y, process='cpu'):
        out = ct.c_float(0.0)
        r = getattr(le.load(), "harmonic_compute_potential_2d_" + process)(ct.byref(self), x, y, ct.byref(out))
        return r, out.value

    def compute_gradient(self, x, y, cdPrecision, process='cpu'):
        px, py = ct.c_float(0.0), ct.c_float(0.0)
        r = getattr(le.load(), "harmonic_compute_gradient_2d_" + process)(ct.byref(self), x, y, cdPrecision,
                                                                          ct.byref(px), ct.byref(py))
        return r, px.value, py.value

    def compute_path(self, x, y, stepSize=0.2, cdPrecision=0.4, maxLength=1000000, process='cpu'):
        """(return code, float32 array of shape (k, 2)); the reference's defaults (harmonic_map.py:117-121)."""
        k = ct.c_uint(0)
        raw = ct.POINTER(ct.c_float)()
        r = getattr(le.load(), "harmonic_compute_path_2d_" + process)(ct.byref(self), x, y, stepSize, cdPrecision,
                                                                      int(maxLength), ct.byref(k), ct.byref(raw))
        if r != le.EPIC_SUCCESS:
            return r, np.zeros((0, 2), np.float32)
        path = np.ctypeslib.as_array(raw, shape=(2 * k.value,)).copy().reshape(-1, 2)
        le.load().harmonic_free_path_cpu(ct.byref(raw))
        return r, path

    def compute_path_poses(self, x, y, stepSize, cdPrecision, maxLength, originX, originY, resolution, process='cpu'):
        """(return code, float32 array (k, 3)): world x, world y, yaw per path point -- what the ROS callers publish."""
        k = ct.c_uint(0)
        raw = ct.POINTER(ct.c_float)()
        r = getattr(le.load(), "harmonic_compute_path_poses_2d_" + process)(
            ct.byref(self), x, y, stepSize, cdPrecision, int(maxLength), originX, originY, resolution, ct.byref(k), ct.byref(raw))
        if r != le.EPIC_SUCCESS:
            return r, np.zeros((0, 3), np.float32)
        out = np.ctypeslib.as_array(raw, shape=(3 * k.value,)).copy().reshape(-1, 3)
        le.load().harmonic_free_path_cpu(ct.byref(raw))
        return r, out

    def compute_paths_gpu(self, starts, stepSize=0.2, cdPrecision=0.4, maxLength=1000000):
        """Many streamlines in one call on the device-resident field: list of (code, path array)."""
        starts = np.ascontiguousarray(starts, dtype=np.float32).reshape(-1, 2)
        n = len(starts)
        rets = (ct.c_int * n)()
        ks = (ct.c_uint * n)()
        raws = (ct.POINTER(ct.c_float) * n)()
        r = le.load().harmonic_compute_paths_2d_gpu(ct.byref(self), n, starts.ctypes.data_as(ct.POINTER(ct.c_float)),
                                                    stepSize, cdPrecision, int(maxLength), rets, ks, raws)
        if r != le.EPIC_SUCCESS:
            raise EpicError("harmonic_compute_paths_2d_gpu", r)
        out = []
        for i in range(n):
            if rets[i] == le.EPIC_SUCCESS:
                out.append((0, np.ctypeslib.as_array(raws[i], shape=(2 * ks[i],)).copy().reshape(-1, 2)))
                le.load().epic_b200_free_path(raws[i])
            else:
                out.append((rets[i], np.zeros((0, 2), np.float32)))
        return out


class HarmonicMap(Harmonic):
    """A 2-D harmonic function loaded from a grayscale PNG (reference harmonic_map.py:62-100)."""

    def __init__(self, filename=None, image=None, **kw):
        Harmonic.__init__(self, **kw)
        if filename is not None:
            self.set_grid(*grids.load_png(filename))
        elif image is not None:
            self.set_grid(*grids.grid_from_image(image))
