#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== gpu tests"; python -m pytest tests -q -m gpu 2>&1 | tail -40
} > gpurun_out/r2i.log 2>&1
tail -60 gpurun_out/r2i.log
