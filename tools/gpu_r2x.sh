#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== conv tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "conv_tc or coupling_fused or sign_bits" 2>&1 | tail -5
echo "== parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -4
echo "== shapes"; python tools/step_shapes.py 2>&1 | grep -E "eager|conv3x3"
echo "== bench"; python bench.py --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k: round(v,3) for k,v in d['profile_ms_per_step'].items()}); print(d.get('inference_1080p'))"
} > gpurun_out/r2x.log 2>&1
tail -40 gpurun_out/r2x.log
