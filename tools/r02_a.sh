#!/bin/bash
# round 2, GPU call A: parity of the new strict tables, A/B against the round-1 library, reference GPU baseline, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== selftest + parity"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -6 | tee gpurun_out/r02a_pytest.log
echo "== A/B strict"; 
for lib in epic_b200/lib/ab/libepic_r1.so epic_b200/lib/libepic.so; do
  echo $lib; EPIC_B200_LIB=$PWD/$lib CONFIGS=512:96,256:64 timeout 600 python tools/sweep_timing.py 16384 strict 400 30 2>&1 | tail -3
done | tee gpurun_out/r02a_ab.log
echo "== reference GPU"
for m in maze umass; do timeout 300 python -m oracle.ref_gpu complete --map $m 2>&1 | tail -1; done | tee gpurun_out/r02a_refgpu.log
timeout 600 python -m oracle.ref_gpu sweeps --size 16384 --steps 20 --warmup 3 2>&1 | tail -1 | tee -a gpurun_out/r02a_refgpu.log
echo "== ours on the same maps"; timeout 300 python tools/solve_timing.py 2>&1 | tail -8 | tee gpurun_out/r02a_solve.log
echo "== ncu strict"; bash tools/gpu_ncu.sh r02a strict
