#!/bin/bash
cd "$(dirname "$0")/.."
for shape in "64 52 52 256 128" "64 52 52 256 256" "64 26 26 424 256" "64 13 13 512 256"; do
  for rep in 1 2 4 8; do for na in 4 2; do
    echo -n "WREP=$rep NA=$na  "; YR_PW_WREP=$rep YR_PW_STREAM_NA=$na timeout 120 python scripts/run_pw_layer.py $shape 3 6 | tail -1
  done; done
done
