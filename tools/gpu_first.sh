#!/bin/bash
# first contact with the GPU: parity tests, keep going after failures to see everything
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
