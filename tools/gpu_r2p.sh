#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== fused coupling kernel tests"; python -m pytest tests/test_gpu_kernels.py -q -x -k "coupling_fused" 2>&1 | tail -8
echo "== parity tests"; python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -4
for m in 5 1 7 0; do echo "== bench 3x3=5 1x1=$m"; SININN_FUSE_COUPLING_1X1=$m python bench.py --no-cpu-baseline --no-inference --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k: round(v,3) for k,v in d['profile_ms_per_step'].items()})"; done
} > gpurun_out/r2p.log 2>&1
tail -40 gpurun_out/r2p.log
