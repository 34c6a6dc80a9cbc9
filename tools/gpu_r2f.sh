#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== gpu tests"; python -m pytest tests -q -x -m gpu 2>&1 | tail -4
echo "== bench"; python bench.py --no-cpu-baseline --no-inference 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['profile_ms_per_step'])"
echo "== timeline eager PDL off"; SININN_PDL=0 MODE=eager python tools/step_timeline.py 2>&1 | grep -v -i warn | head -14
} > gpurun_out/r2f.log 2>&1
tail -40 gpurun_out/r2f.log
