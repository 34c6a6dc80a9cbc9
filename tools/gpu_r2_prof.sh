#!/bin/bash
# profile captures at HEAD: eager-step launch list + --set full of the dominant kernels (one ncu process family per call)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2b}
CMD="python bench.py --no-graph --steps 2 --warmup 3 --no-extras --no-inference --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
for c in l1_conv2 c3x3_conv2 c3x3_conv1 wgg_l1_3x3 wgg_l0_3x3 bwd1x1 s1x1_fwd coupling_bwd; do
  python tools/ncu_one.py $c > gpurun_out/${TAG}_${c}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"conv_tc_pair|wgrad_pair|subnet1x1|coupling_bwd" -s 2 -c 1 -f -o gpurun_out/${TAG}_${c} python tools/ncu_one.py $c > gpurun_out/${TAG}_${c}_ncu.log 2>&1
done
ls -la gpurun_out | grep ${TAG} | head -40
tail -3 gpurun_out/${TAG}_ncu_launch.log
