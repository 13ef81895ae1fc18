#!/bin/bash
# the CTA-pair kernel (variant 4) on the wide head layers: CTA-0 timelines and one ncu --set full capture each
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for shape in "64 52 52 256 128" "64 52 52 128 256" "64 52 52 128 384"; do
  n=$(echo $shape | tr ' ' '_')
  YR_PW_TC_DEBUG=1 timeout 120 python scripts/run_pw_layer.py $shape 4 1 2>&1 | grep -A9 "timeline" | head -11 > gpurun_out/r2_tl_pair_$n.log
  timeout 120 python scripts/run_pw_layer.py $shape 4 8 2>/dev/null | tail -1
  timeout 120 python scripts/run_pw_layer.py $shape 3 8 2>/dev/null | tail -1
done
for shape in "64 52 52 256 128" "64 52 52 128 256"; do
  n=$(echo $shape | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_ts_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_pair_$n -f python scripts/run_pw_layer.py $shape 4 2 > /dev/null 2>&1
done
ls gpurun_out | grep pair
