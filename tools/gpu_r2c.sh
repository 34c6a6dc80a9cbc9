#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest wgrad"; python -m pytest tests/test_gpu_kernels.py -q -x -k "wgrad" 2>&1 | tail -5
echo "== bench pair + TMA-store epilogue"; KT=0 ONLY=wgrad python tools/bench_kernels.py 2>&1 | tail -12
echo "== PDL off, kernel times"; SININN_PDL=0 ONLY=wgrad python tools/bench_kernels.py 2>&1 | grep -v Warn | tail -12
echo "== trace"; python tools/wgrad_trace.py 2>&1 | tail -10
echo "== timeline graph"; python tools/step_timeline.py 2>&1 | tail -45
echo "== timeline eager PDL off"; SININN_PDL=0 MODE=eager python tools/step_timeline.py 2>&1 | tail -45
} > gpurun_out/r2c.log 2>&1
tail -150 gpurun_out/r2c.log
