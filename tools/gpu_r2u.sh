#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "subnet1x1_fused_backward or subnet1x1_backward" 2>&1 | tail -5
timeout 120 python tools/s1bwd_trace.py
timeout 120 python tools/ncu_one.py bwd1x1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none python tools/ncu_one.py bwd1x1 2>&1 | grep -E "subnet1x1_bwd_kernel|wgrad_reduce_kernel|gpu__time" | tail -4
} > gpurun_out/r2u.log 2>&1
tail -40 gpurun_out/r2u.log
