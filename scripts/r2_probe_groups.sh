#!/bin/bash
# GPU tests, two bench lines, then the stride-1 fused layers alone (event-timed, L2 flushed) with 2 / 3 converter groups
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']), d['verified'], {k:v for k,v in d['pw_kernel_picks'].items() if k!='per_layer'}, {k:v['ms_per_step'] for k,v in d['kernels'].items() if k in ('pw','dwpw')})"; done
for shape in "64 104 104 144 24 1" "64 52 52 144 24 1" "64 26 26 288 48 1" "64 26 26 432 72 1" "64 13 13 720 120 1" "64 13 13 720 240 1"; do for g in 2 3; do echo -n "G=$g "; YR_DWPW_GROUPS=$g YR_ONLY_FUSED=1 timeout 120 python scripts/run_dwpw_layer.py $shape 8 2>/dev/null | tail -1; done; done
