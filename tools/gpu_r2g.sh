#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== gpu tests"; python -m pytest tests -q -x -m gpu 2>&1 | tail -25
echo "== bench"; python bench.py --no-cpu-baseline --no-inference 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['profile_ms_per_step'])"
} > gpurun_out/r2g.log 2>&1
tail -60 gpurun_out/r2g.log
