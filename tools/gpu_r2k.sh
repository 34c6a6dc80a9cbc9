#!/bin/bash
# two-GPU run: data-parallel step with the forward half's all-reduce hidden behind the inverse half
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== gpu tests (device guard test needs 2 GPUs)"; python -m pytest tests -q -m gpu -k "non_current_device or overlapped or graph_replay" 2>&1 | tail -5
echo "== bench N=2"; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/r2k_n2.err | tail -1 > gpurun_out/r2k_n2.json; tail -3 gpurun_out/r2k_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_n2.json'))
print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'], d['inference_1080p']['fwd_inv_frames_per_s'], d['inference_1080p']['e2e_120_frames'])
PY
echo "== bench N=1"; python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'])"
echo "== DP equivalence: 2 ranks x batch 4 vs 1 rank x batch 8"; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/dp_check.py 2>&1 | tail -4
} > gpurun_out/r2k.log 2>&1
tail -40 gpurun_out/r2k.log
