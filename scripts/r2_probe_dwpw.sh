#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for shape in "64 104 104 144 24 1" "64 208 208 24 16 1" "64 26 26 432 72 1" "64 208 208 96 24 2" "64 13 13 720 120 1"; do
  tag=$(echo $shape | tr ' ' '_')
  YR_ONLY_FUSED=1 YR_PW_TC_DEBUG=1 timeout 120 python scripts/run_dwpw_layer.py $shape 2 > gpurun_out/r2_dwtl_$tag.log 2>&1
  timeout 120 python scripts/run_dwpw_layer.py $shape 5 | tail -1
done
YR_ONLY_FUSED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_ts_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_dwpw_block2 -f python scripts/run_dwpw_layer.py 64 104 104 144 24 1 2 > /dev/null 2>&1
ls -la gpurun_out | grep ncu_dwpw
