#!/bin/bash
# NCCL CTA-count sweep at N GPUs.  Every run is wrapped in `timeout`: on 2026-10-18 a multi-rank run without one hung at exit
# (graph-captured NCCL work alive at destroy_process_group, fixed in bench.py since) and consumed the round's remaining GPU budget.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
{
for CT in default 2 4 8 16; do
  echo "== N=$N NCCL_MAX_CTAS=$CT"
  if [ "$CT" = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$CT; fi
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 --no-inference --no-extras --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['ms_per_step_spread'])"
done
} > gpurun_out/r2dp$N.log 2>&1
tail -12 gpurun_out/r2dp$N.log
