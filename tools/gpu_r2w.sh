#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
MODE=graph python tools/step_timeline.py 2>&1 | tail -40
} > gpurun_out/r2w.log 2>&1
tail -50 gpurun_out/r2w.log
