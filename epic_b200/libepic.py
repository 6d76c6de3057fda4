"""ctypes bindings of epic_b200/lib/libepic.so.

Mirror of the reference's libepic/python/epic/epic_harmonic.py:38-124: the same `EpicHarmonic`
structure (field for field) and the same `argtypes` for the 30 reference entry points, plus this
library's extensions (device-side streamlines, include/epic/libepic.h; slab API, include/epic_b200.h).
"""
import ctypes as ct
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EPIC_B200_LIB") or os.path.join(HERE, "lib", "libepic.so")   # override: A/B runs of two builds

EPIC_SUCCESS = 0
EPIC_SUCCESS_AND_CONVERGED = 1
EPIC_ERROR_INVALID_DATA = 2
EPIC_ERROR_INVALID_CUDA_PARAM = 3
EPIC_ERROR_DEVICE_MALLOC = 4
EPIC_ERROR_MEMCPY_TO_DEVICE = 5
EPIC_ERROR_MEMCPY_TO_HOST = 6
EPIC_ERROR_DEVICE_FREE = 7
EPIC_ERROR_KERNEL_EXECUTION = 8
EPIC_ERROR_DEVICE_SYNCHRONIZE = 9
EPIC_ERROR_INVALID_LOCATION = 10
EPIC_ERROR_INVALID_CELL_TYPE = 11
EPIC_ERROR_INVALID_GRADIENT = 12
EPIC_ERROR_INVALID_PATH = 13


class EpicHarmonic(ct.Structure):
    """The C struct Harmonic (include/epic/libepic.h; reference harmonic.h:44-64)."""

    _fields_ = [("n", ct.c_uint),
                ("m", ct.POINTER(ct.c_uint)),
                ("u", ct.POINTER(ct.c_float)),
                ("locked", ct.POINTER(ct.c_uint)),
                ("epsilon", ct.c_float),
                ("delta", ct.c_float),
                ("numIterationsToStaggerCheck", ct.c_uint),
                ("currentIteration", ct.c_uint),
                ("d_m", ct.POINTER(ct.c_uint)),
                ("d_u", ct.POINTER(ct.c_float)),
                ("d_locked", ct.POINTER(ct.c_uint)),
                ("d_delta", ct.POINTER(ct.c_float))]


class FieldInfo(ct.Structure):
    _fields_ = [("pitch", ct.c_uint64), ("layer_floats", ct.c_uint64), ("launches", ct.c_uint64),
                ("device_bytes", ct.c_uint64), ("sweeps_per_pass", ct.c_uint32), ("tile_rows", ct.c_uint32),
                ("math", ct.c_uint32), ("device", ct.c_int32), ("skipped_tiles", ct.c_uint64)]


class GridStats(ct.Structure):
    """epic_b200_stats (include/epic_b200.h)."""
    _fields_ = [("slabs", ct.c_uint32), ("last_solve_iterations", ct.c_uint32), ("last_solve_delta", ct.c_float),
                ("reserved", ct.c_uint32), ("last_solve_seconds", ct.c_double), ("launches", ct.c_uint64),
                ("skipped_tiles", ct.c_uint64 * 16)]


def build(force=False):
    """Compile libepic.so in-tree (nvcc, sm_100a)."""
    if force:
        subprocess.run(["make", "-C", os.path.join(HERE, "csrc"), "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libepic.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


_P = ct.POINTER
_H = _P(EpicHarmonic)

# name -> argtypes; every entry returns int.  (References are pointers at the ABI.)
REFERENCE_EXPORTS = {
    "harmonic_complete_cpu": (_H,),
    "harmonic_update_cpu": (_H,),
    "harmonic_update_and_check_cpu": (_H,),
    "harmonic_complete_gpu": (_H, ct.c_uint),
    "harmonic_initialize_gpu": (_H, ct.c_uint),
    "harmonic_execute_gpu": (_H, ct.c_uint),
    "harmonic_uninitialize_gpu": (_H,),
    "harmonic_update_gpu": (_H, ct.c_uint),
    "harmonic_update_and_check_gpu": (_H, ct.c_uint),
    "harmonic_get_potential_values_gpu": (_H,),
    "harmonic_initialize_dimension_size_gpu": (_H,),
    "harmonic_uninitialize_dimension_size_gpu": (_H,),
    "harmonic_initialize_potential_values_gpu": (_H,),
    "harmonic_uninitialize_potential_values_gpu": (_H,),
    "harmonic_initialize_locked_gpu": (_H,),
    "harmonic_uninitialize_locked_gpu": (_H,),
    "harmonic_update_model_gpu": (_H,),
    "harmonic_utilities_set_cells_2d_cpu": (_H, ct.c_uint, _P(ct.c_uint), _P(ct.c_uint)),
    "harmonic_utilities_set_cells_2d_gpu": (_H, ct.c_uint, ct.c_uint, _P(ct.c_uint), _P(ct.c_uint)),
    "harmonic_compute_potential_2d_cpu": (_H, ct.c_float, ct.c_float, _P(ct.c_float)),
    "harmonic_compute_gradient_2d_cpu": (_H, ct.c_float, ct.c_float, ct.c_float, _P(ct.c_float), _P(ct.c_float)),
    "harmonic_compute_path_2d_cpu": (_H, ct.c_float, ct.c_float, ct.c_float, ct.c_float, ct.c_uint,
                                     _P(ct.c_uint), _P(_P(ct.c_float))),
    "harmonic_free_path_cpu": (_P(_P(ct.c_float)),),
    "harmonic_legacy_sor_2d_float_cpu": (ct.c_uint, ct.c_uint, ct.c_float, ct.c_float, _P(ct.c_uint),
                                         _P(ct.c_float), _P(ct.c_uint)),
    "harmonic_legacy_sor_2d_double_cpu": (ct.c_uint, ct.c_uint, ct.c_double, ct.c_double, _P(ct.c_uint),
                                          _P(ct.c_double), _P(ct.c_uint)),
    "harmonic_legacy_sor_2d_long_double_cpu": (ct.c_uint, ct.c_uint, ct.c_longdouble, ct.c_longdouble,
                                               _P(ct.c_uint), _P(ct.c_longdouble), _P(ct.c_uint)),
    "harmonic_legacy_compute_potential_2d_cpu": (ct.c_uint, ct.c_uint, _P(ct.c_uint), _P(ct.c_double), ct.c_double,
                                                 ct.c_double, _P(ct.c_double)),
    "harmonic_legacy_compute_gradient_2d_cpu": (ct.c_uint, ct.c_uint, _P(ct.c_uint), _P(ct.c_double), ct.c_double,
                                                ct.c_double, ct.c_double, _P(ct.c_double), _P(ct.c_double)),
    "harmonic_legacy_compute_path_2d_cpu": (ct.c_uint, ct.c_uint, _P(ct.c_uint), _P(ct.c_double), ct.c_double,
                                            ct.c_double, ct.c_double, ct.c_double, ct.c_uint, ct.c_int,
                                            _P(ct.c_uint), _P(_P(ct.c_double))),
    "harmonic_legacy_free_path_cpu": (_P(_P(ct.c_double)),),
}

_F = ct.c_void_p
EXTENSION_EXPORTS = {
    "harmonic_legacy_sor_2d_float_gpu": (ct.c_uint, ct.c_uint, ct.c_float, ct.c_float, _P(ct.c_uint),
                                         _P(ct.c_float), _P(ct.c_uint)),
    "harmonic_legacy_sor_2d_double_gpu": (ct.c_uint, ct.c_uint, ct.c_double, ct.c_double, _P(ct.c_uint),
                                          _P(ct.c_double), _P(ct.c_uint)),
    "harmonic_compute_potential_2d_gpu": (_H, ct.c_float, ct.c_float, _P(ct.c_float)),
    "harmonic_compute_gradient_2d_gpu": (_H, ct.c_float, ct.c_float, ct.c_float, _P(ct.c_float), _P(ct.c_float)),
    "harmonic_compute_path_2d_gpu": (_H, ct.c_float, ct.c_float, ct.c_float, ct.c_float, ct.c_uint,
                                     _P(ct.c_uint), _P(_P(ct.c_float))),
    "harmonic_compute_paths_2d_gpu": (_H, ct.c_uint, _P(ct.c_float), ct.c_float, ct.c_float, ct.c_uint,
                                      _P(ct.c_int), _P(ct.c_uint), _P(_P(ct.c_float))),
    "harmonic_path_to_poses_2d": (_P(ct.c_float), ct.c_uint, ct.c_float, ct.c_float, ct.c_float, _P(ct.c_float)),
    "harmonic_compute_path_poses_2d_cpu": (_H, ct.c_float, ct.c_float, ct.c_float, ct.c_float, ct.c_uint, ct.c_float,
                                           ct.c_float, ct.c_float, _P(ct.c_uint), _P(_P(ct.c_float))),
    "harmonic_compute_path_poses_2d_gpu": (_H, ct.c_float, ct.c_float, ct.c_float, ct.c_float, ct.c_uint, ct.c_float,
                                           ct.c_float, ct.c_float, _P(ct.c_uint), _P(_P(ct.c_float))),
    "harmonic_utilities_set_occupancy_grid_2d_cpu": (_H, _P(ct.c_byte), ct.c_int, ct.c_int),
    "harmonic_utilities_set_occupancy_grid_2d_gpu": (_H, _P(ct.c_byte), ct.c_int, ct.c_int),
    "harmonic_utilities_reset_free_cells_2d_cpu": (_H,),
    "harmonic_utilities_reset_free_cells_2d_gpu": (_H,),
    "epic_b200_field_create": (_P(_F), ct.c_uint, _P(ct.c_uint64), ct.c_uint64, ct.c_uint64, ct.c_uint, ct.c_int,
                               ct.c_int, ct.c_void_p, ct.c_int),
    "epic_b200_field_info": (_F, _P(FieldInfo)),
    "epic_b200_field_upload_u": (_F, _P(ct.c_float), ct.c_uint64, ct.c_uint64),
    "epic_b200_field_upload_locked": (_F, _P(ct.c_uint32), ct.c_uint64, ct.c_uint64),
    "epic_b200_field_download_u": (_F, _P(ct.c_float), ct.c_uint64, ct.c_uint64),
    "epic_b200_field_download_locked": (_F, _P(ct.c_uint32), ct.c_uint64, ct.c_uint64),
    "epic_b200_field_run": (_F, ct.c_uint32, ct.c_uint32, ct.c_int),
    "epic_b200_field_read_delta": (_F, _P(ct.c_float)),
    "epic_b200_field_solve": (_F, ct.c_float, ct.c_uint32, ct.c_uint32, _P(ct.c_uint32), _P(ct.c_float)),
    "epic_b200_field_sync": (_F,),
    "epic_b200_field_set_tracking": (_F, ct.c_int),
    "epic_b200_field_set_cells_2d": (_F, ct.c_uint32, _P(ct.c_uint32), _P(ct.c_uint32)),
    "epic_b200_field_potential_2d": (_F, ct.c_float, ct.c_float, _P(ct.c_float)),
    "epic_b200_field_gradient_2d": (_F, ct.c_float, ct.c_float, ct.c_float, _P(ct.c_float), _P(ct.c_float)),
    "epic_b200_field_peer_export": (_F, ct.c_void_p, ct.c_uint64),
    "epic_b200_field_set_peer_ipc": (_F, ct.c_int, ct.c_void_p, ct.c_uint64),
    "epic_b200_field_set_peer_local": (_F, ct.c_int, _F),
    "epic_b200_harmonic_stats": (ct.c_void_p, _P(GridStats)),
    "epic_b200_selftest_math": (ct.c_uint32, _P(ct.c_uint64), _P(ct.c_uint64), _P(ct.c_uint64), _P(ct.c_uint64)),
    "epic_b200_field_paths_2d": (_F, ct.c_uint32, _P(ct.c_float), ct.c_float, ct.c_float, ct.c_uint32, _P(ct.c_int),
                                 _P(ct.c_uint32), _P(_P(ct.c_float))),
}
# entries that do not return int
_SPECIAL = {
    "epic_b200_field_destroy": ((_F,), None),
    "epic_b200_field_layer_ptr": ((_F, ct.c_int64), ct.c_void_p),
    "epic_b200_free_path": ((_P(ct.c_float),), None),
    "epic_b200_version": ((), ct.c_char_p),
}
ALL_EXPORTS = sorted(list(REFERENCE_EXPORTS) + list(EXTENSION_EXPORTS) + list(_SPECIAL))

_lib = None


def load():
    """The loaded library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run epic_b200.libepic.build() (or python -c 'import __graft_entry__ as g; "
                               "g.build()') first" % LIB_PATH)
        lib = ct.CDLL(LIB_PATH)
        for table in (REFERENCE_EXPORTS, EXTENSION_EXPORTS):
            for name, argtypes in table.items():
                try:
                    fn = getattr(lib, name)
                except AttributeError:
                    if os.environ.get("EPIC_B200_LIB"):
                        continue    # an older build loaded for an A/B timing run
                    raise
                fn.argtypes = argtypes
                fn.restype = ct.c_int
        for name, (argtypes, restype) in _SPECIAL.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = lib
    return _lib
