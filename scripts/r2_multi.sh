#!/bin/bash
# usage: r2_multi.sh N  -> bench lines of the four multi-GPU configs at N ranks (BASELINE.json configs[1..4])
cd "$(dirname "$0")/.."
N=$1
mkdir -p gpurun_out
for w in cfg2 cfg3 cfg4 cfg5; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 --workload $w --no-cpu-baseline --no-roofline > gpurun_out/r2_n${N}_$w.log 2> gpurun_out/r2_n${N}_$w.err
  echo "== $w rc=$?"; grep -h '^{' gpurun_out/r2_n${N}_$w.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    e = d.get('e2e') or {}
    print(d['n_gpus'], round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(e.get('value', 0)), 'f32-upload', round((e.get('fp32_upload') or {}).get('value', 0)), 'verified', d.get('verified'), d.get('multi_gpu_equivalence'))
"; tail -3 gpurun_out/r2_n${N}_$w.err | cut -c1-300
done
