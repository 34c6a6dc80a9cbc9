#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for V in 0 1; do
  echo "== SININN_FUSE_STORE=$V"
  SININN_FUSE_STORE=$V python bench.py --no-cpu-baseline --no-inference --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k: round(v,3) for k,v in d['profile_ms_per_step'].items()})"
done
} > gpurun_out/r2y.log 2>&1
tail -50 gpurun_out/r2y.log
