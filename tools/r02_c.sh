#!/bin/bash
# round 2, GPU call C: 3 CTAs/SM strict candidates, branch-free streamline step, sanitizer, default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -5 | tee gpurun_out/r02c_pytest.log
echo "== streamlines"; timeout 300 python tools/path_timing.py 2>&1 | tail -6 | tee gpurun_out/r02c_paths.log
echo "== strict tile candidates (256 threads: 3 CTAs/SM when the tile allows)"
CONFIGS=256:64,256:56,256:48,256:40,256:72,512:96 timeout 900 python tools/sweep_timing.py 16384 strict 400 30 2>&1 | tail -7 | tee gpurun_out/r02c_tiles.log
CONFIGS=256:48,256:64,512:64 timeout 600 python tools/sweep_timing.py 16384 fast 400 30 2>&1 | tail -4 | tee -a gpurun_out/r02c_tiles.log
echo "== bench default"; timeout 1200 python bench.py > gpurun_out/r02c_bench_default.json 2> gpurun_out/r02c_bench_default.err; tail -c 1500 gpurun_out/r02c_bench_default.json; tail -3 gpurun_out/r02c_bench_default.err
echo "== sanitizer"; bash tools/r02_sanitizer.sh
