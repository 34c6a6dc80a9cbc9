#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== IRN shapes"; ARCH=IRN python tools/step_shapes.py 2>&1 | head -8
echo "== bench IRN"; python bench.py --no-cpu-baseline --no-inference 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); print(d['irn_arch']); print(d['deep_variant']['value'], d['fp32_path']['value'])"
} > gpurun_out/r2y.log 2>&1
tail -40 gpurun_out/r2y.log
