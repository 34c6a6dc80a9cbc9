#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== bench full"; SECONDS=0; python bench.py 2> gpurun_out/r2j_bench.err | tail -1 > gpurun_out/r2j_bench.json; echo "bench wall ${SECONDS}s"; tail -5 gpurun_out/r2j_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
for k in ('value','ms_per_step','e2e','gpu_launches','clocks','fp32_path','fp32_cuda_core_path','deep_variant','irn_arch','config0_cpu','cpu_baseline'):
    print(k, d.get(k))
r=d['roofline']; print({k:r[k] for k in r if k not in ('families','timing_note','traffic_note')})
for k,v in r['families'].items(): print(' ',k,v)
print(d['roofline_hbm'])
print(d.get('inference_1080p'))
PY
echo "== reference arm"; python bench.py --impl reference --steps 3 --warmup 1 | tail -1 | cut -c1-400
} > gpurun_out/r2j.log 2>&1
tail -60 gpurun_out/r2j.log
