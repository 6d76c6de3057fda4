#!/bin/bash
# final 8-GPU evidence: 16384^2 bench at N=8 (both modes, time-to-epsilon) and the 1024^3 x0-slab run
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/gpu_multi.sh "8" 5
python -c "
import json
d=json.loads(open('gpurun_out/scale_n8.json').read().strip().splitlines()[-1]); print(json.dumps(d['time_to_epsilon'])[:600])"
for m in fast strict; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29755 \
      tools/sharded3d_timing.py 1024 $m 3 2>&1 | grep -E "3-D|Error|error" | tail -3
done | tee gpurun_out/sharded3d_timing.log
