#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest wgrad"; python -m pytest tests/test_gpu_kernels.py -q -x -k "wgrad" 2>&1 | tail -5
echo "== PDL off, kernel times"; SININN_PDL=0 ONLY=wgrad python tools/bench_kernels.py 2>&1 | grep -v -i Warn | sed 's/cudaLaunchKernel=0.0us  Activity Buffer Request=0.0us  cudaLaunchKernelExC=0.0us//; s/cudaDeviceSynchronize=0.0us//' | tail -12
echo "== trace"; python tools/wgrad_trace.py 2>&1 | tail -10
echo "== full gpu tests"; python -m pytest tests -q -x -m gpu 2>&1 | tail -5
echo "== bench"; python bench.py --no-cpu-baseline 2>&1 | tail -3
} > gpurun_out/r2d.log 2>&1
tail -60 gpurun_out/r2d.log
