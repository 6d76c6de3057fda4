#!/bin/bash
# round 2, GPU call G (1 GPU): tests after the streamline rewrite and the legacy SOR kernels; streamline timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -6 | tee gpurun_out/r02g_pytest.log
echo "== streamlines"; timeout 300 python tools/path_timing.py 2>&1 | tail -6 | tee gpurun_out/r02g_paths.log
echo "== legacy SOR timing"; timeout 600 python tools/legacy_timing.py 2>&1 | tail -6 | tee gpurun_out/r02g_legacy.log
