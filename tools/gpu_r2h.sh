#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== gpu tests"; python -m pytest tests -q -m gpu 2>&1 | tail -25
echo "== bench bf16"; python bench.py --no-cpu-baseline --no-inference 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
echo "== bench fp32 (tensor-core split)"; python bench.py --precision fp32tc --no-cpu-baseline --no-inference --steps 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['profile_ms_per_step'])"
echo "== bench fp32 (CUDA cores) B=8"; python bench.py --precision fp32 --batch 8 --no-cpu-baseline --no-inference --steps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
} > gpurun_out/r2h.log 2>&1
tail -60 gpurun_out/r2h.log
