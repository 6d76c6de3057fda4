#!/bin/bash
# round 2, GPU call D (2 GPUs): Grid over two real devices, NCCL / IPC sharding, max-size sharded test, bench at N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -3
echo "== multi-GPU tests"; timeout 1500 python -m pytest tests/test_grid_gpu.py tests/test_sharded_gpu.py -m gpu -q -x --timeout 1000 2>&1 | tail -8 | tee gpurun_out/r02d_pytest.log
echo "== bench N=2 (2-D)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 5 --warmup 3 --no-tte-all-tiles > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err
tail -c 2500 gpurun_out/r02d_bench_n2.json; tail -5 gpurun_out/r02d_bench_n2.err
echo "== bench N=2 (3-D 512^3)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --dims 3 --size 512 --steps 3 --warmup 3 --no-tte-all-tiles --no-abi-multi > gpurun_out/r02d_bench3d_n2.json 2> gpurun_out/r02d_bench3d_n2.err
tail -c 1200 gpurun_out/r02d_bench3d_n2.json; tail -5 gpurun_out/r02d_bench3d_n2.err
