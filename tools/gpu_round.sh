#!/bin/bash
# whole-round evidence on one B200: GPU parity suite, then bench + ncu (tools/gpu_bench.sh)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
bash tools/gpu_bench.sh
