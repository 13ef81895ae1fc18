#!/bin/bash
# column-split epilogue (both groups drain every tile) vs alternating tiles: GPU tests, bench, N-heavy layers alone
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
for sp in 0 1; do
echo "== YR_PW_EPI_SPLIT=$sp"
for shape in "64 26 26 48 288" "64 26 26 72 432" "64 13 13 120 720" "64 52 52 24 144" "64 104 104 24 144" "64 52 52 128 256"; do YR_PW_EPI_SPLIT=$sp timeout 120 python scripts/run_pw_layer.py $shape 3 8 2>/dev/null | tail -1; done
YR_PW_EPI_SPLIT=$sp YR_ONLY_FUSED=1 timeout 120 python scripts/run_dwpw_layer.py 64 208 208 24 96 1 8 2>/dev/null | tail -1
YR_PW_EPI_SPLIT=$sp timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']), d['verified'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if k in ('pw','dwpw')})"
done
