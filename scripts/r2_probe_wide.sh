#!/bin/bash
# timelines (CTA 0) + ncu full captures of the wide head pointwise shapes (round 2 probe)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for shape in "64 52 52 256 128" "64 52 52 128 256" "64 52 52 256 256" "64 26 26 424 256" "64 104 104 144 24" "64 26 26 72 432"; do
  tag=$(echo $shape | tr ' ' '_')
  YR_PW_TC_DEBUG=1 python scripts/run_pw_layer.py $shape 3 3 > gpurun_out/r2_tl_$tag.log 2>&1
  python scripts/run_pw_layer.py $shape 3 5 | tail -1
  python scripts/run_pw_layer.py $shape 2 5 | tail -1
done
for shape in "64 52 52 256 128" "64 52 52 128 256" "64 52 52 256 256"; do
  tag=$(echo $shape | tr ' ' '_')
  ncu --set full --clock-control none --import-source on -k regex:pw_ts_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_$tag -f python scripts/run_pw_layer.py $shape 3 3 > /dev/null 2>&1
done
ls -la gpurun_out | grep r2_ncu
