"""The drop-in boundary, checked without a GPU: the library loads, exports every symbol the headers
declare (and therefore everything the reference's Python wrapper binds at import,
libepic/python/epic/epic_harmonic.py:61-124), keeps the 80-byte Harmonic layout, returns the
reference's error codes for bad input, and its *_cpu entry points reproduce the reference CPU results
bit for bit (golden vectors made with the untouched reference code)."""
import ctypes as ct
import os
import re
import subprocess

import numpy as np
import pytest

import common
from epic_b200 import libepic as le
from epic_b200.harmonic import Harmonic

ROOT = common.ROOT


def declared_symbols():
    names = set()
    for header in ("include/epic/libepic.h", "include/epic_b200.h"):
        text = open(os.path.join(ROOT, header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b((?:harmonic|epic_b200)_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol(libepic_built):
    out = subprocess.run(["nm", "-D", "--defined-only", le.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    declared = declared_symbols()
    assert len(declared) == 30 + 4 + 3 + 4 + 23 + 1 + 2   # reference, device streamlines, pose lists, dense map ingest, slab API, stats, legacy SOR on the GPU
    assert declared <= exported, "missing exports: %s" % sorted(declared - exported)
    assert set(le.ALL_EXPORTS) == declared, "Python binding table and headers disagree"
    assert len(le.REFERENCE_EXPORTS) == 30
    for name in le.ALL_EXPORTS:
        getattr(libepic_built, name)


def test_library_has_no_cuda_shared_object_dependencies(libepic_built):
    """libcudart is linked statically and libcuda is resolved at run time, so the reference's callers
    link with -lepic alone (CMakeLists.txt:66-75 of the reference) and the library loads without a driver."""
    out = subprocess.run(["ldd", le.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "libcuda" not in out and "libcudart" not in out


def test_library_contains_sm100a_code_only(libepic_built):
    out = subprocess.run(["cuobjdump", "--list-elf", le.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_struct_layout_matches_reference():
    assert ct.sizeof(le.EpicHarmonic) == 80
    offsets = {name: getattr(le.EpicHarmonic, name).offset for name, _ in le.EpicHarmonic._fields_}
    assert offsets == {"n": 0, "m": 8, "u": 16, "locked": 24, "epsilon": 32, "delta": 36,
                       "numIterationsToStaggerCheck": 40, "currentIteration": 44, "d_m": 48, "d_u": 56,
                       "d_locked": 64, "d_delta": 72}


def test_header_compiles_as_c_and_cpp(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include <epic/libepic.h>\n#include <epic_b200.h>\n'
                   'int main(void){ Harmonic h; (void)h; return sizeof(Harmonic) == 80 ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(tmp_path / "tc")], check=True)
    subprocess.run([str(tmp_path / "tc")], check=True)
    cpp = tmp_path / "t.cpp"
    # the include set of the reference's ROS nodes (src/epic_navigation_node_harmonic.cpp:33-40)
    cpp.write_text("".join("#include <epic/%s>\n" % h for h in (
        "harmonic/harmonic_cpu.h", "harmonic/harmonic_gpu.h", "harmonic/harmonic_model_gpu.h",
        "harmonic/harmonic_path_cpu.h", "harmonic/harmonic_utilities_cpu.h", "harmonic/harmonic_utilities_gpu.h",
        "error_codes.h", "constants.h", "harmonic/harmonic.h")) +
        "int main(){ epic::Harmonic h; h.d_u = nullptr; (void)h.d_u; unsigned int k = 0; float *p = nullptr;\n"
        " int (*f)(epic::Harmonic*, float, float, float, float, unsigned int, unsigned int&, float*&) = "
        "epic::harmonic_compute_path_2d_cpu; (void)f; (void)k; (void)p; return EPIC_SUCCESS; }\n")
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-I", inc, "-c", str(cpp), "-o", str(tmp_path / "t.o")],
                   check=True)


# ---- error behaviour that needs no device ---------------------------------------------------------

def test_gpu_entry_points_reject_bad_input_with_reference_codes(libepic_built):
    L = libepic_built
    h = Harmonic()
    assert L.harmonic_initialize_dimension_size_gpu(ct.byref(h)) == le.EPIC_ERROR_INVALID_DATA   # n == 0
    assert L.harmonic_initialize_potential_values_gpu(ct.byref(h)) == le.EPIC_ERROR_INVALID_DATA
    assert L.harmonic_initialize_locked_gpu(ct.byref(h)) == le.EPIC_ERROR_INVALID_DATA
    assert L.harmonic_initialize_gpu(ct.byref(h), 1024) == le.EPIC_ERROR_INVALID_DATA
    assert L.harmonic_execute_gpu(ct.byref(h), 1024) == le.EPIC_ERROR_INVALID_DATA
    assert L.harmonic_update_model_gpu(ct.byref(h)) == le.EPIC_ERROR_INVALID_DATA
    assert L.harmonic_update_gpu(ct.byref(h), 1024) == le.EPIC_ERROR_INVALID_DATA
    assert L.harmonic_get_potential_values_gpu(ct.byref(h)) == le.EPIC_ERROR_INVALID_DATA
    # uninitialize of nothing is a success and leaves null handles (harmonic_model_gpu.cu:60-72)
    assert L.harmonic_uninitialize_dimension_size_gpu(ct.byref(h)) == 0
    assert L.harmonic_uninitialize_potential_values_gpu(ct.byref(h)) == 0
    assert L.harmonic_uninitialize_locked_gpu(ct.byref(h)) == 0
    assert L.harmonic_uninitialize_gpu(ct.byref(h)) == 0
    assert not h.d_m and not h.d_u and not h.d_locked and not h.d_delta
    u, locked, _, _ = common.case_input("box64")
    h = Harmonic(u, locked, 1e-3, 100)
    v = np.zeros(2, np.uint32)
    assert h.set_cells(v[:0], v[:0], "cpu") if False else True
    assert L.harmonic_utilities_set_cells_2d_gpu(ct.byref(h), 1024, 0, v.ctypes.data_as(ct.POINTER(ct.c_uint)),
                                                 v.ctypes.data_as(ct.POINTER(ct.c_uint))) == le.EPIC_ERROR_INVALID_DATA
    # a stale / foreign device handle is recognised, not dereferenced
    h.d_m = ct.cast(ct.c_void_p(0xdead0000), ct.POINTER(ct.c_uint))
    h.d_u = ct.cast(ct.c_void_p(0xdead0000), ct.POINTER(ct.c_float))
    h.d_locked = ct.cast(ct.c_void_p(0xdead0000), ct.POINTER(ct.c_uint))
    assert L.harmonic_execute_gpu(ct.byref(h), 1024) == le.EPIC_ERROR_INVALID_DATA
    assert L.harmonic_execute_gpu(ct.byref(h), 1000) == le.EPIC_ERROR_INVALID_CUDA_PARAM


def test_dimension_handle_lifecycle_without_device(libepic_built):
    """initialize_dimension_size / initialize_gpu only create host-side state, as far as a caller can tell."""
    L = libepic_built
    u, locked, _, _ = common.case_input("box64")
    h = Harmonic(u, locked, 1e-3, 100)
    assert L.harmonic_initialize_dimension_size_gpu(ct.byref(h)) == 0 and h.d_m
    assert L.harmonic_initialize_gpu(ct.byref(h), 1024) == 0 and h.d_delta
    assert L.harmonic_initialize_gpu(ct.byref(h), 1024) == le.EPIC_ERROR_INVALID_DATA   # d_delta already set
    assert L.harmonic_uninitialize_gpu(ct.byref(h)) == 0 and not h.d_delta
    assert L.harmonic_uninitialize_dimension_size_gpu(ct.byref(h)) == 0 and not h.d_m


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="checks the no-GPU behaviour")
def test_gpu_solve_fails_loudly_without_a_gpu(libepic_built):
    u, locked, _, _ = common.case_input("box64")
    h = Harmonic(u, locked, 1e-3, 100)
    r = libepic_built.harmonic_complete_gpu(ct.byref(h), 1024)
    assert r in (le.EPIC_ERROR_DEVICE_MALLOC, le.EPIC_ERROR_INVALID_CUDA_PARAM), "no silent CPU fallback: got %d" % r
    assert np.array_equal(h.field, u), "host field must be untouched when the GPU path fails"


# ---- the library's CPU exports against the reference golden vectors ----------------------------------

def make_cpu(u, locked, eps, stagger):
    return common.LibepicSolver(u, locked, eps, stagger, "cpu")


@pytest.mark.parametrize("name", ["box64", "random256", "random_ragged", "random48x3", "random3d_ragged",
                                  "proc_maze", "basic", "umass", "maze"])
def test_cpu_exports_checkpoints(golden, libepic_built, name):
    common.check_checkpoints(make_cpu, name, golden[name])


@pytest.mark.parametrize("name", ["box64", "random_ragged", "random3d_ragged", "proc_maze"])
def test_cpu_exports_full_solve_and_paths(golden, libepic_built, name):
    s = common.check_complete(make_cpu, name, golden[name])
    common.check_paths(s, golden[name])
    common.check_potentials(s, golden[name])


def test_cpu_exports_set_cells(golden, libepic_built):
    s = common.check_set_cells(make_cpu, golden["set_cells"])
    assert common.sha1(s.h.locked_cells) == golden["set_cells"]["sha1_locked_final"]


def test_cpu_complete_rejects_bad_input(libepic_built):
    u, locked, _, _ = common.case_input("box64")
    for eps in (0.0, -1.0):
        h = Harmonic(u.copy(), locked.copy(), eps, 100)
        assert libepic_built.harmonic_complete_cpu(ct.byref(h)) == le.EPIC_ERROR_INVALID_DATA
    h = Harmonic()
    assert libepic_built.harmonic_complete_cpu(ct.byref(h)) == le.EPIC_ERROR_INVALID_DATA


def test_path_entry_requires_null_path_and_frees(libepic_built, golden):
    u, locked, eps, stagger = common.case_input("box64")
    s = make_cpu(u, locked, eps, stagger)
    s.complete()
    k = ct.c_uint(0)
    raw = ct.POINTER(ct.c_float)()
    L = libepic_built
    assert L.harmonic_compute_path_2d_cpu(ct.byref(s.h), 50.0, 50.0, 0.05, 0.5, 81920, ct.byref(k), ct.byref(raw)) == 0
    assert k.value == golden["box64"]["paths"][0]["k"] and raw
    # path != nullptr on entry is an error (harmonic_path_cpu.cpp:158-164)
    assert L.harmonic_compute_path_2d_cpu(ct.byref(s.h), 50.0, 50.0, 0.05, 0.5, 81920, ct.byref(k),
                                          ct.byref(raw)) == le.EPIC_ERROR_INVALID_DATA
    assert L.harmonic_free_path_cpu(ct.byref(raw)) == 0 and not raw
    assert L.harmonic_free_path_cpu(ct.byref(raw)) == 0


def test_legacy_sor_exports(libepic_built):
    """Linear-space SOR (out of the GPU scope, exported for the Python wrapper): a 1-D-like ramp converges
    to the harmonic solution; all three precisions agree."""
    w, h = 12, 5
    locked = np.ones((h, w), np.uint32)
    locked[1:-1, 1:-1] = 0
    res = {}
    for name, dt, ctype in (("float", np.float32, ct.c_float), ("double", np.float64, ct.c_double),
                            ("long_double", np.longdouble, ct.c_longdouble)):
        u = np.zeros((h, w), dt)
        u[:, 0] = 1.0
        u[0, :] = u[-1, :] = np.linspace(1.0, 0.0, w)
        it = ct.c_uint(0)
        fn = getattr(libepic_built, "harmonic_legacy_sor_2d_%s_cpu" % name)
        assert fn(w, h, ctype(1e-6), ctype(1.5), locked.ctypes.data_as(ct.POINTER(ct.c_uint)),
                  u.ctypes.data_as(ct.POINTER(ctype)), ct.byref(it)) == 0
        assert it.value >= 10000
        res[name] = u.astype(np.float64)
        assert np.allclose(res[name][2], np.linspace(1.0, 0.0, w), atol=1e-4)
    assert np.allclose(res["float"], res["double"], atol=1e-4)
    assert np.allclose(res["double"], res["long_double"], atol=1e-9)


def test_slab_info_struct_matches_python_mirror(tmp_path):
    """epic_b200_info (include/epic_b200.h) and its ctypes mirror must agree field by field."""
    src = tmp_path / "info.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "epic_b200.h"\nint main(void){\n'
                   + "".join('printf("%s %%zu\\n", offsetof(epic_b200_info, %s));\n' % (n, n) for n, _ in le.FieldInfo._fields_)
                   + 'printf("sizeof %zu\\n", sizeof(epic_b200_info)); return 0; }\n')
    exe = tmp_path / "info"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(common.ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, _ in le.FieldInfo._fields_:
        assert int(out[name]) == getattr(le.FieldInfo, name).offset, name
    assert int(out["sizeof"]) == ct.sizeof(le.FieldInfo)
