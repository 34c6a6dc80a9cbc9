#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== kernel tests"; python -m pytest tests/test_gpu_kernels.py -q -x -k coupling_fused 2>&1 | tail -12
echo "== parity tests"; python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -30
for m in "5 1" "0 0" "5 0" "1 1"; do set -- $m; echo "== bench store 3x3=$1 1x1=$2"; SININN_FUSE_COUPLING=$1 SININN_FUSE_COUPLING_1X1=$2 python bench.py --no-cpu-baseline --no-inference --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k: round(v,3) for k,v in d['profile_ms_per_step'].items()}, d.get('peak_mem_gb'))"; done
} > gpurun_out/r2q.log 2>&1
tail -60 gpurun_out/r2q.log
