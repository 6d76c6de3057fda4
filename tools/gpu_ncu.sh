#!/bin/bash
# ncu --set full of the sweep kernel; usage: gpu_ncu.sh tag math [size]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; m=$2; size=${3:-16384}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep -s 100 -c 1 -f -o gpurun_out/${tag}_$m \
    python tools/profile_sweep.py $m $size 400 6 > gpurun_out/ncu_${tag}_$m.log 2>&1
tail -2 gpurun_out/ncu_${tag}_$m.log
