#!/bin/bash
# round-2 call A: pair weight-gradient kernel -- correctness at both halo widths, timing vs the round-1 kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest wgrad, halo 10"; python -m pytest tests/test_gpu_kernels.py -q -x -k "wgrad" 2>&1 | tail -15
echo "== pytest wgrad, halo 16"; SININN_WG_HALO=16 python -m pytest tests/test_gpu_kernels.py -q -x -k "wgrad" 2>&1 | tail -15
echo "== bench pair halo 10"; ONLY=wgrad python tools/bench_kernels.py 2>&1 | tail -12
echo "== bench pair halo 16"; SININN_WG_HALO=16 ONLY=wgrad python tools/bench_kernels.py 2>&1 | tail -12
echo "== bench round-1 kernel"; SININN_WG_PAIR=0 ONLY=wgrad python tools/bench_kernels.py 2>&1 | tail -12
echo "== full gpu tests"; python -m pytest tests -q -x -m gpu 2>&1 | tail -15
} > gpurun_out/r2a.log 2>&1
tail -80 gpurun_out/r2a.log
