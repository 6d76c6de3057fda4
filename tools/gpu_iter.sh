#!/bin/bash
# quick iteration: parity subset + timing variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sharded_gpu.py -m gpu -q -x --timeout 600 -k "not full_size" 2>&1 | tail -8
if [ -n "$TIMING" ]; then
for m in strict fast; do CONFIGS=${CONFIGS:-512:96,256:48} timeout 600 python tools/sweep_timing.py 16384 $m 400 30 2>&1 | tail -12; done | tee gpurun_out/timing.log
fi
if [ -n "$MISC" ]; then timeout 900 python tools/misc_timing.py 2>&1 | tail -12 | tee gpurun_out/misc_timing.log; fi
