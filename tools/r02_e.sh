#!/bin/bash
# round 2, GPU call E (8 GPUs): BASELINE configs 3, 5 and 4 at N = 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { # name, port, args...
  name=$1; port=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "== $name rc=$?"; tail -c 900 gpurun_out/$name.json | head -c 900; echo; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/$name.err | tail -4
}
run r02e_bench_n8 29511 --steps 10 --warmup 3 --no-tte-all-tiles
run r02e_bench3d_n8 29512 --dims 3 --size 1024 --steps 5 --warmup 3 --no-tte-all-tiles --no-abi-multi
run r02e_maze65536_n8 29513 --workload maze --size 65536 --steps 3 --warmup 3 --no-tte --no-abi-multi
run r02e_maze65536_tte_n8 29514 --workload maze --size 65536 --corridor 2046 --goals 16 --math fast --single-mode --steps 3 --warmup 3 --no-tte-all-tiles --no-abi-multi
