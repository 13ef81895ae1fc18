#!/bin/bash
# ncu --set full (with SASS-level sampling) of the fused depthwise->pointwise kernel on block_2's shape, 2 and 3 converter groups
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for g in 2 3; do
YR_DWPW_GROUPS=$g YR_ONLY_FUSED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_ts_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_dwpw_block2_g$g -f python scripts/run_dwpw_layer.py 64 104 104 144 24 1 2 > /dev/null 2>&1
done
YR_DWPW_GROUPS=3 YR_ONLY_FUSED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_ts_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_dwpw_26_g3 -f python scripts/run_dwpw_layer.py 64 26 26 432 72 1 2 > /dev/null 2>&1
ls -la gpurun_out | grep ncu_dwpw
