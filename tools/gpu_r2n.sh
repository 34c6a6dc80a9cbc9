#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --no-cpu-baseline --no-inference --no-extras 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); print(d['profile_ms_per_step'])
r=d['roofline']
for k,v in r['families'].items(): print(k, round(v['achieved']), round(v['frac'],3), round(v['ms_per_step'],3))
for k,v in d['roofline_hbm'].items(): print(k, round(v['achieved']), round(v['frac'],3), round(v['ms_per_step'],3))
print('alg only', r['algorithmic_only_frac'], 'whole', r['whole_step_frac'], 'sum', sum(d['profile_ms_per_step'].values()))"
