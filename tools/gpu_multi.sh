#!/bin/bash
# multi-GPU: NCCL/P2P parity test + bench at the given GPU counts.  usage: gpu_multi.sh "2 4 8" [steps] [extra bench args]
cd "$(dirname "$0")/.."
NS=${1:-2}; STEPS=${2:-10}; EXTRA=${3:-}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
if [ -z "$SKIP_NCCL_TEST" ]; then timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q -x --timeout 500 -k nccl 2>&1 | tail -3; fi
for N in $NS; do
  if [ "$N" = "1" ]; then
    timeout 400 python bench.py --gpus 1 --steps $STEPS --warmup 3 --no-cpu $EXTRA > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N \
      bench.py --gpus $N --steps $STEPS --warmup 3 --no-cpu $EXTRA > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_n$N.json").read().strip().splitlines()[-1])
    o=d.get("fast_mode") or d.get("strict_mode")
    print("N=$N strict %.1f GCUPS (%.2f ms/step, kern %.3f, e2e %.1f, TH %s) | fast %.1f GCUPS (%.2f ms/step, kern %.3f, e2e %.1f)"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["e2e"]["value"],d["config"]["tile_rows"],o["value"],o["ms_per_step"],o["roofline"]["kernel_ms"],o["e2e"]["value"]))
except Exception as e:
    print("N=$N failed", e); print(open("gpurun_out/scale_n$N.err").read()[-1500:])
PY
done
