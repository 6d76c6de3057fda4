#!/bin/bash
# ncu evidence for the current kernels: launch list of a short bench run + --set full captures of both modes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-tte > gpurun_out/bench_under_ncu.log 2>&1
for m in strict fast; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep2d -s 100 -c 2 -f -o gpurun_out/sweep2d_$m \
      python tools/profile_sweep.py $m 16384 400 6 > gpurun_out/ncu_$m.log 2>&1
  tail -1 gpurun_out/ncu_$m.log
done
