"""epic_b200/csrc/kernels/strict_math.h (the bit-exact twins of glibc expf / logf that the CUDA sweep
replays) against THIS host's libm, exhaustively: every float x <= 0 for expf (2.1e9 values) and every float
in [1/8, 16) for logf.  The header compiles as plain C++ for this purpose; the GPU twin of this test is
tests/test_parity_gpu.py::test_strict_math_device_functions_equal_host_libm."""
import os
import subprocess

import common


def test_strict_math_header_equals_host_libm_exhaustively(tmp_path):
    exe = str(tmp_path / "strict_math_check")
    src = os.path.join(common.ROOT, "tests", "native", "strict_math_check.cpp")
    subprocess.run(["g++", "-O2", "-fopenmp", "-ffp-contract=off", src, "-o", exe, "-lm"], check=True)
    r = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    assert "expf mismatches 0 of 2139095041" in r.stdout and "logf mismatches 0 of" in r.stdout
    assert "update4 mismatches 0 of 40000000" in r.stdout
