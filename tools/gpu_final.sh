#!/bin/bash
# final evidence of the round: full GPU test suite, default bench (+ reference arm), profiles at HEAD
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== gpu tests"; SECONDS=0; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4; echo "tests wall ${SECONDS}s"
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
} > gpurun_out/final_tests.log 2>&1
bash tools/gpu_r2r.sh r2final > /dev/null 2>&1
{ echo "== reference arm"; python bench.py --impl reference --steps 3 --warmup 1 | tail -1 | cut -c1-600; } >> gpurun_out/r2final.log 2>&1
bash tools/gpu_r2_prof.sh r2c > gpurun_out/r2c_prof.log 2>&1
tail -8 gpurun_out/final_tests.log
grep -E "^value|^ms_per_step |^e2e|^irn_arch|^inference" gpurun_out/r2final.log | cut -c1-300
