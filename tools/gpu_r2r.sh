#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== gpu tests"; SECONDS=0; python -m pytest tests -m gpu -x -q 2>&1 | tail -8; echo "tests wall ${SECONDS}s"
echo "== bench full"; SECONDS=0; python bench.py 2> gpurun_out/r2r_bench.err | tail -1 > gpurun_out/r2r_bench.json; echo "bench wall ${SECONDS}s"; tail -5 gpurun_out/r2r_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2r_bench.json'))
for k in d:
    if k not in ('roofline','roofline_hbm'): print(k, d[k])
r=d['roofline']; print({k:r[k] for k in r if k not in ('families','timing_note','traffic_note')})
for k,v in r['families'].items(): print(' ',k,v)
print(d['roofline_hbm'])
PY
} > gpurun_out/r2r.log 2>&1
tail -80 gpurun_out/r2r.log
