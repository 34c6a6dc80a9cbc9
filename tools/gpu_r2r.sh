#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2r}
{
echo "== bench full"; SECONDS=0; python bench.py 2> gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json; echo "bench wall ${SECONDS}s"; tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
for k in d:
    if k not in ('roofline','roofline_hbm','config'): print(k, d[k])
r=d['roofline']; print({k:r[k] for k in r if k not in ('families','timing_note','traffic_note')})
for k,v in r['families'].items(): print(' ',k,v)
print(d['roofline_hbm'])
PY
ARCH=IRN python tools/step_shapes.py 2>&1 | head -30
} > gpurun_out/${TAG}.log 2>&1
tail -90 gpurun_out/${TAG}.log
