#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== conv tests"; python -m pytest tests/test_gpu_kernels.py -q -x -k "conv_tc" 2>&1 | tail -3
echo "== kbench acc4"; python tools/bench_kernels.py 2>&1 | grep -E "3x3" 
echo "== kbench acc2"; SININN_PAIR_ACC4=0 python tools/bench_kernels.py 2>&1 | grep -E "3x3"
echo "== bench"; python bench.py --no-cpu-baseline --no-inference --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['profile_ms_per_step'])"
} > gpurun_out/r2l.log 2>&1
tail -40 gpurun_out/r2l.log
