#!/bin/bash
# round 2, GPU call F (1 GPU): BASELINE config 5 at N = 1 (bench line + ncu at 1024^3, not 512^3), streamline kernel profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== bench 3-D 1024^3 N=1"
timeout 1500 python bench.py --dims 3 --size 1024 --steps 5 --warmup 3 --no-tte-all-tiles > gpurun_out/r02f_bench3d_n1.json 2> gpurun_out/r02f_bench3d_n1.err
tail -c 600 gpurun_out/r02f_bench3d_n1.json; tail -3 gpurun_out/r02f_bench3d_n1.err
echo "== ncu 3-D at 1024^3"
for m in fast strict; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep3d -s 60 -c 1 -f -o gpurun_out/r02f3d_$m \
    python tools/profile_sweep.py $m 1024x1024x1024 120 4 > gpurun_out/ncu_r02f3d_$m.log 2>&1; tail -2 gpurun_out/ncu_r02f3d_$m.log
done
echo "== ncu streamline kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:path_2d -c 1 -f -o gpurun_out/r02f_path python tools/path_timing.py > gpurun_out/ncu_r02f_path.log 2>&1; tail -3 gpurun_out/ncu_r02f_path.log
