#!/bin/bash
# N-GPU bench with the two orderings of passes between GPUs (in-kernel flags vs stream memory operations).  usage: gpu_p2p_modes.sh N [steps]
cd "$(dirname "$0")/.."
N=${1:-2}; STEPS=${2:-5}
mkdir -p gpurun_out
for mode in kernel stream kernel stream; do
  EPIC_P2P_SYNC=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N \
      bench.py --gpus $N --steps $STEPS --warmup 3 --no-cpu --no-tte > gpurun_out/p2p_${mode}_n$N.json 2> gpurun_out/p2p_${mode}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/p2p_${mode}_n$N.json").read().strip().splitlines()[-1]); o=d["fast_mode"]
    print("N=$N sync=$mode strict %.1f GCUPS (kern %.4f ms) | fast %.1f GCUPS (kern %.4f ms)"%(d["value"],d["roofline"]["kernel_ms"],o["value"],o["roofline"]["kernel_ms"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/p2p_${mode}_n$N.err").read()[-1500:])
PY
done
