#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== bench pair halo 10"; ONLY=wgrad python tools/bench_kernels.py 2>&1 | tail -12
echo "== bench round-1 kernel"; SININN_WG_PAIR=0 ONLY=wgrad python tools/bench_kernels.py 2>&1 | tail -12
} > gpurun_out/r2b.log 2>&1
tail -80 gpurun_out/r2b.log
