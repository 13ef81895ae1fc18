#!/bin/bash
# ncu --set full of the plain tcgen05 pointwise kernel on two N-heavy backbone expand layers (per-role sampling input)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_ts_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_expand_26 -f python scripts/run_pw_layer.py 64 26 26 48 288 3 2 > gpurun_out/r2_expand_26.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_ts_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_expand_52 -f python scripts/run_pw_layer.py 64 52 52 24 144 3 2 > gpurun_out/r2_expand_52.log 2>&1
python scripts/run_pw_layer.py 64 26 26 48 288 3 6 2>&1 | tail -2
python scripts/run_pw_layer.py 64 26 26 48 288 4 6 2>&1 | tail -2
ls gpurun_out | grep expand
