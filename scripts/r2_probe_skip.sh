#!/bin/bash
# Where does the fused depthwise front spend its time?  Event-timed runs of one layer with parts of the kernel switched
# off (debug instantiation, YR_PW_TC_DEBUG=2 = no timeline / no sync).  YR_DWPW_SKIP bits: 1 = no depthwise FMAs,
# 2 = no TMA box loads, 4 = no staging / split / TMEM stores, 8 = no MMAs.
cd "$(dirname "$0")/.."
run() { echo -n "$1 groups=$2 skip=$3  "; YR_DWPW_GROUPS=$2 YR_DWPW_SKIP=$3 YR_ONLY_FUSED=1 YR_PW_TC_DEBUG=2 timeout 120 python scripts/run_dwpw_layer.py $1 6 2>/dev/null | tail -1; }
for g in 2 3; do for skip in 0 1 2 4 5 7 15; do run "64 104 104 144 24 1" $g $skip; done; done
for skip in 0 2 5 7; do run "64 26 26 432 72 1" 3 $skip; done
