"""The slab object of include/epic_b200.h from Python (ctypes): one x0-range of a grid resident on one
B200, with optional ghost layers for neighbouring ranks."""
import ctypes as ct

import numpy as np

from . import libepic as le

MATH = {"strict": 0, "fast": 1, "env": -1}


class DevicePointer:
    """A device address dressed up with __cuda_array_interface__ so torch.as_tensor can alias it."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class Field:
    def __init__(self, shape, row0=0, rows=None, ghost=0, math="strict", device=-1, stream=None):
        self.shape = tuple(int(s) for s in shape)
        self.row0 = int(row0)
        self.rows = self.shape[0] - self.row0 if rows is None else int(rows)
        self.ghost = int(ghost)
        self._lib = le.load()
        m = (ct.c_uint64 * len(self.shape))(*self.shape)
        h = ct.c_void_p()
        r = self._lib.epic_b200_field_create(ct.byref(h), len(self.shape), m, self.row0, self.rows, self.ghost,
                                             MATH[math], int(device), ct.c_void_p(stream or 0),
                                             0 if stream is None else 1)
        if r != 0:
            raise RuntimeError("epic_b200_field_create failed with libepic error %d" % r)
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.epic_b200_field_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, name, r):
        if r != 0:
            raise RuntimeError("%s failed with libepic error %d" % (name, r))

    def info(self):
        i = le.FieldInfo()
        self._check("info", self._lib.epic_b200_field_info(self._h, ct.byref(i)))
        return {k: getattr(i, k) for k, _ in le.FieldInfo._fields_}

    def _layers(self, first, layers):
        first = self.row0 if first is None else first
        layers = self.rows if layers is None else layers
        return int(first), int(layers)

    def upload(self, u, locked, first=None, layers=None):
        first, layers = self._layers(first, layers)
        u = np.ascontiguousarray(u, dtype=np.float32)
        locked = np.ascontiguousarray(locked, dtype=np.uint32)
        assert u.shape == (layers,) + self.shape[1:] and locked.shape == u.shape
        self._check("upload_u", self._lib.epic_b200_field_upload_u(
            self._h, u.ctypes.data_as(ct.POINTER(ct.c_float)), first, layers))
        self._check("upload_locked", self._lib.epic_b200_field_upload_locked(
            self._h, locked.ctypes.data_as(ct.POINTER(ct.c_uint32)), first, layers))

    def upload_u(self, u, first=None, layers=None):
        first, layers = self._layers(first, layers)
        u = np.ascontiguousarray(u, dtype=np.float32)
        if u.shape != (layers,) + self.shape[1:]:     # the C side would read past a short buffer
            raise ValueError("upload_u: array of shape %r does not cover %d layers of %r" % (u.shape, layers, self.shape[1:]))
        self._check("upload_u", self._lib.epic_b200_field_upload_u(
            self._h, u.ctypes.data_as(ct.POINTER(ct.c_float)), first, layers))

    def download_u(self, first=None, layers=None, out=None):
        first, layers = self._layers(first, layers)
        if out is None:
            out = np.empty((layers,) + self.shape[1:], np.float32)
        elif out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"] or out.size != layers * int(np.prod(self.shape[1:])):
            raise ValueError("download_u: `out` must be a C-contiguous float32 array of %d x %r values" % (layers, self.shape[1:]))
        self._check("download_u", self._lib.epic_b200_field_download_u(
            self._h, out.ctypes.data_as(ct.POINTER(ct.c_float)), first, layers))
        return out

    def download_locked(self, first=None, layers=None):
        first, layers = self._layers(first, layers)
        out = np.empty((layers,) + self.shape[1:], np.uint32)
        self._check("download_locked", self._lib.epic_b200_field_download_locked(
            self._h, out.ctypes.data_as(ct.POINTER(ct.c_uint32)), first, layers))
        return out

    def run(self, it0, count, check_last=False):
        self._check("run", self._lib.epic_b200_field_run(self._h, int(it0), int(count), 1 if check_last else 0))

    def read_delta(self):
        d = ct.c_float(0.0)
        self._check("read_delta", self._lib.epic_b200_field_read_delta(self._h, ct.byref(d)))
        return d.value

    def solve(self, epsilon=1e-3, stagger=100, m_max=None):
        it, d = ct.c_uint32(0), ct.c_float(0.0)
        m_max = max(self.shape) if m_max is None else m_max
        self._check("solve", self._lib.epic_b200_field_solve(self._h, epsilon, stagger, m_max, ct.byref(it),
                                                             ct.byref(d)))
        return it.value, d.value

    def set_tracking(self, on):
        """Static-tile skipping for the passes run() issues from now on (bit-identical; see epic_b200.h)."""
        self._check("set_tracking", self._lib.epic_b200_field_set_tracking(self._h, 1 if on else 0))

    def sync(self):
        self._check("sync", self._lib.epic_b200_field_sync(self._h))

    def layer_ptr(self, layer):
        return self._lib.epic_b200_field_layer_ptr(self._h, int(layer))

    PEER_BLOB_BYTES = 256

    def peer_export(self):
        """Opaque bytes (CUDA IPC handles + geometry) a neighbouring rank needs for peer-to-peer halos."""
        buf = ct.create_string_buffer(self.PEER_BLOB_BYTES)
        self._check("peer_export", self._lib.epic_b200_field_peer_export(self._h, buf, self.PEER_BLOB_BYTES))
        return buf.raw

    def set_peer_ipc(self, direction, blob):
        buf = ct.create_string_buffer(bytes(blob), self.PEER_BLOB_BYTES)
        return self._lib.epic_b200_field_set_peer_ipc(self._h, int(direction), buf, self.PEER_BLOB_BYTES)

    def set_peer_local(self, direction, other):
        return self._lib.epic_b200_field_set_peer_local(self._h, int(direction), other._h)

    def set_cells(self, v, types):
        v = np.ascontiguousarray(v, dtype=np.uint32).reshape(-1)
        types = np.ascontiguousarray(types, dtype=np.uint32)
        return self._lib.epic_b200_field_set_cells_2d(self._h, len(types), v.ctypes.data_as(ct.POINTER(ct.c_uint32)),
                                                      types.ctypes.data_as(ct.POINTER(ct.c_uint32)))

    def paths(self, starts, step=0.05, cd=0.5, max_length=1000000):
        starts = np.ascontiguousarray(starts, dtype=np.float32).reshape(-1, 2)
        n = len(starts)
        rets, ks, raws = (ct.c_int * n)(), (ct.c_uint32 * n)(), (ct.POINTER(ct.c_float) * n)()
        self._check("paths", self._lib.epic_b200_field_paths_2d(
            self._h, n, starts.ctypes.data_as(ct.POINTER(ct.c_float)), step, cd, int(max_length), rets, ks, raws))
        out = []
        for i in range(n):
            if rets[i] == 0:
                out.append((0, np.ctypeslib.as_array(raws[i], shape=(2 * ks[i],)).copy().reshape(-1, 2)))
                self._lib.epic_b200_free_path(raws[i])
            else:
                out.append((rets[i], np.zeros((0, 2), np.float32)))
        return out
