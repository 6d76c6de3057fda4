#!/bin/bash
# round 2, GPU call B: Grid behind the ABI, warp-cooperative streamlines, new golden cases, tile candidates, reference GPU (nobar)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -12 | tee gpurun_out/r02b_pytest.log
echo "== streamlines"; timeout 300 python tools/path_timing.py 2>&1 | tail -6 | tee gpurun_out/r02b_paths.log
echo "== strict tile candidates"
CONFIGS=256:96,256:80,256:64,256:56,256:48,512:96,512:80,512:64 timeout 900 python tools/sweep_timing.py 16384 strict 400 30 2>&1 | tail -9 | tee gpurun_out/r02b_tiles.log
echo "== reference GPU"
(timeout 25 python -m oracle.ref_gpu complete --map basic 2>&1 | tail -1; echo "stock rc=$?") | tee gpurun_out/r02b_refgpu.log
timeout 300 python -m oracle.ref_gpu sweeps --variant nobar --size 16384 --steps 10 --warmup 2 2>&1 | tail -1 | tee -a gpurun_out/r02b_refgpu.log
for m in maze umass; do it=49301; [ $m = umass ] && it=32701; timeout 200 python -m oracle.ref_gpu updates --variant nobar --map $m --iterations $it 2>&1 | tail -1; done | tee -a gpurun_out/r02b_refgpu.log
