#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
python tools/step_shapes.py
ARCH=IRN python tools/step_shapes.py
} > gpurun_out/r2s.log 2>&1
tail -150 gpurun_out/r2s.log
