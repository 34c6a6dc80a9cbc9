#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -x 2>&1 | tail -5
echo "== shapes"; python tools/step_shapes.py 2>&1 | grep -E "eager|layout|resample|permute|coupling"
echo "== bench"; python bench.py --no-cpu-baseline --no-inference --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k: round(v,3) for k,v in d['profile_ms_per_step'].items()})"
} > gpurun_out/r2v.log 2>&1
tail -30 gpurun_out/r2v.log
