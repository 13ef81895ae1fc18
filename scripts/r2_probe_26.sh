cd /root/repo
for nb in 2 4; do echo -n "MIN_NB=$nb "; YR_DWPW_MIN_NB=$nb YR_ONLY_FUSED=1 timeout 120 python scripts/run_dwpw_layer.py 64 26 26 432 72 1 8 2>/dev/null | tail -1; done
for nb in 2 4; do echo -n "MIN_NB=$nb "; YR_DWPW_MIN_NB=$nb YR_ONLY_FUSED=1 timeout 120 python scripts/run_dwpw_layer.py 64 26 26 288 48 1 8 2>/dev/null | tail -1; done
YR_ONLY_FUSED=1 YR_PW_TC_DEBUG=1 timeout 120 python scripts/run_dwpw_layer.py 64 26 26 432 72 1 1 2>&1 | grep -A9 "dwpw timeline" | head -12 > gpurun_out/r2_dwtl_26_432_72.log
YR_ONLY_FUSED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_ts_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_dwpw_26_final -f python scripts/run_dwpw_layer.py 64 26 26 432 72 1 2 > /dev/null 2>&1
YR_ONLY_FUSED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_ts_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_dwpw_block2_final -f python scripts/run_dwpw_layer.py 64 104 104 144 24 1 2 > /dev/null 2>&1
ls gpurun_out | grep final
