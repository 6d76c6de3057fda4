"""The reference's OWN Python wrapper (libepic/python/epic/{epic_harmonic,harmonic,harmonic_map}.py), unmodified and
imported from where it lies under /root/reference, bound to this repository's libepic.so: every `argtypes`
assignment must find its symbol at import (epic_harmonic.py:61-124), `HarmonicMap.load` + `Harmonic.solve` must run,
and the results must be the golden ones.  Container only (the reference tree does not travel to the GPU box).

Two adaptations, both outside the reference's files: the wrapper looks for `../../lib/libepic.so` next to its own
(read-only) source, so ctypes.CDLL is redirected to the drop-in for that one path; and `time.clock`, which
harmonic.py:80 still calls, no longer exists in Python >= 3.8 (a reference bug noted in SURVEY.md section 2.1), so
the test supplies it."""
import ctypes as ct
import importlib
import os
import sys
import time

import numpy as np
import pytest

import common
from epic_b200 import libepic as le
from oracle import oracle as orc

REF_PY = "/root/reference/libepic/python/epic"
REF_MAPS = "/root/reference/libepic/tests/maps"

pytestmark = pytest.mark.skipif(not os.path.isdir(REF_PY), reason="the reference tree is only present in the build container")


@pytest.fixture(scope="module")
def ref_wrapper(libepic_built):
    real = ct.CDLL
    loaded = []

    def redirected(path, *a, **k):
        if isinstance(path, str) and path.endswith(os.path.join("lib", "libepic.so")) and "reference" in os.path.realpath(path):
            loaded.append(path)
            path = le.LIB_PATH
        return real(path, *a, **k)
    had_clock = hasattr(time, "clock")
    ct.CDLL = redirected
    if not had_clock:
        time.clock = time.process_time
    sys.path.insert(0, REF_PY)
    try:
        for name in ("epic_harmonic", "harmonic", "harmonic_map"):
            sys.modules.pop(name, None)
        eh = importlib.import_module("epic_harmonic")
        hm = importlib.import_module("harmonic_map")
        assert loaded, "the wrapper did not try to load ../../lib/libepic.so"
        yield eh, hm
    finally:
        ct.CDLL = real
        sys.path.remove(REF_PY)
        if not had_clock:
            del time.clock
        for name in ("epic_harmonic", "harmonic", "harmonic_map", "harmonic_legacy"):
            sys.modules.pop(name, None)


def test_unmodified_wrapper_binds_all_thirty_symbols(ref_wrapper):
    eh, _ = ref_wrapper
    for name in le.REFERENCE_EXPORTS:
        assert getattr(eh._epic, name).argtypes is not None, name
    assert ct.sizeof(eh.EpicHarmonic) == 80


def test_unmodified_wrapper_solves_and_traces_on_the_cpu_exports(ref_wrapper, golden):
    """HarmonicMap.load (the reference's PNG loader) -> Harmonic.solve(process='cpu') -> _compute_streamline."""
    _, hm = ref_wrapper
    m = hm.HarmonicMap()
    m.load(os.path.join(REF_MAPS, "basic.png"))
    wall, cpu = m.solve(process="cpu", epsilon=1e-3)
    assert wall > 0.0
    g = golden["basic"]["complete"]
    assert m.currentIteration == g["iterations"]
    assert common.hexf(m.delta) == g["delta_hex"]
    u = np.ctypeslib.as_array(m.u, shape=(m.m[0] * m.m[1],)).reshape(m.m[0], m.m[1]).copy()
    assert common.sha1(u) == g["sha1_u"]
    # a streamline with the wrapper's own parameters (0.2, 0.4, 1e6) against the oracle on the same field
    u0, locked, eps, stagger = common.case_input("basic")
    o = orc.Oracle(u.copy(), locked, eps, stagger)
    ys, xs = np.nonzero(locked == 0)
    x, y = float(xs[len(xs) // 3]), float(ys[len(ys) // 3])
    want_r, want = o.path(x, y, 0.2, 0.4, 1000000)
    assert want_r == 0
    got = np.array(m._compute_streamline(x, y), np.float32)
    assert np.array_equal(got, want)


def test_unmodified_wrapper_gpu_request_falls_back_like_the_reference(ref_wrapper, capsys):
    """Without a device the *_gpu calls return the reference's error codes and the wrapper's own fallback
    (harmonic.py:73-97) takes the CPU path; with a B200 the same call is the CUDA path."""
    _, hm = ref_wrapper
    m = hm.HarmonicMap()
    m.load(os.path.join(REF_MAPS, "basic.png"))
    m.solve(process="gpu", epsilon=1e-1)
    assert m.currentIteration > 0
    u = np.ctypeslib.as_array(m.u, shape=(m.m[0] * m.m[1],)).reshape(m.m[0], m.m[1]).copy()
    u0, locked, _, _ = common.case_input("basic")
    o = orc.Oracle(u0.copy(), locked.copy(), 1e-1, 100)
    assert o.complete() == 0
    assert m.currentIteration == o.iteration and np.array_equal(u, o.u)
