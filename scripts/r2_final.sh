#!/bin/bash
# round-2 closing evidence on one GPU: tests, the full bench line, the reference arm, aux workloads, the ncu step capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_final_tests.log; tail -2 gpurun_out/r2_final_tests.log
YR_BENCH_LAYERS=gpurun_out/r2_layers_final.json timeout 900 python bench.py --steps 50 --warmup 10 > gpurun_out/r2_bench_n1.log 2> gpurun_out/r2_bench_n1.err; head -c 200 gpurun_out/r2_bench_n1.log; echo
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.log 2>&1; head -c 200 gpurun_out/r2_bench_reference_arm.log; echo
for w in cfg3 cfg4 cfg5 post post0; do timeout 600 python bench.py --workload $w --steps 30 --no-cpu-baseline >> gpurun_out/r2_bench_aux_workloads.log 2>/dev/null; done; cut -c1-160 gpurun_out/r2_bench_aux_workloads.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_step_raw.csv python bench.py --ncu-step > /dev/null 2>&1
python scripts/ncu_step_summary.py gpurun_out/r2_step_raw.csv gpurun_out/r2_step_ncu 64 | head -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
