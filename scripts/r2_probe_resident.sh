cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
for rb in 1 2; do echo -n "RES_BOXES=$rb "; YR_DWPW_RES_BOXES=$rb YR_ONLY_FUSED=1 timeout 120 python scripts/run_dwpw_layer.py 64 26 26 288 48 1 8 2>/dev/null | tail -1; done
for i in 1 2; do timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']), d['verified'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if k in ('pw','dwpw')})"; done
timeout 300 python bench.py --workload cfg3 --steps 30 --no-cpu-baseline 2>/dev/null | cut -c1-150
timeout 300 python bench.py --workload cfg4 --steps 30 --no-cpu-baseline 2>/dev/null | cut -c1-150
