#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
make -C sin_inn_b200/csrc clean > /dev/null; make -C sin_inn_b200/csrc -j16 EXTRA=-DSININN_PAIR_TRACE 2>&1 | grep -E "error" 
python tools/pair_trace.py c1 d2
} > gpurun_out/r2y.log 2>&1
tail -40 gpurun_out/r2y.log
