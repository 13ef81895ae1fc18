#!/bin/bash
cd "$(dirname "$0")/.."
for g in 2 3; do for skip in 0 1 2 3; do
  echo -n "groups=$g skip=$skip  "; YR_DWPW_GROUPS=$g YR_DWPW_SKIP=$skip YR_ONLY_FUSED=1 YR_PW_TC_DEBUG=1 timeout 120 python scripts/run_dwpw_layer.py 64 104 104 144 24 1 4 2>/dev/null | tail -1
done; done
