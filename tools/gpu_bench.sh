#!/bin/bash
# bench + ncu evidence on one B200; everything lands in gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --tte --tte-max-iterations 60000 > gpurun_out/bench_strict.json 2> gpurun_out/bench_strict.err; tail -c 3000 gpurun_out/bench_strict.json; tail -5 gpurun_out/bench_strict.err
timeout 600 python bench.py --steps 5 --warmup 3 --math fast --no-cpu --no-tte > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err; tail -c 3000 gpurun_out/bench_fast.json; tail -5 gpurun_out/bench_fast.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-tte > gpurun_out/bench_under_ncu.log 2>&1
# full capture of the sweep kernel, steady-state field
for m in strict fast; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep2d -s 100 -c 2 -f -o gpurun_out/sweep2d_$m \
      python tools/profile_sweep.py $m 16384 400 6 > gpurun_out/ncu_$m.log 2>&1
  tail -3 gpurun_out/ncu_$m.log
done
ls -la gpurun_out
