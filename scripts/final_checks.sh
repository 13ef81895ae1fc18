#!/bin/bash
# Round-end evidence run on the GPU box (one gpurun call): tests, smoke, sanitizer, aux workloads, ncu launch list, bench.
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/fin_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/fin_smoke.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > $O/fin_sanitizer.log 2>&1; echo "memcheck exit $?" >> $O/fin_sanitizer.log
: > $O/fin_aux.log
for w in cfg3 cfg4 cfg5 post post0; do timeout 600 python bench.py --workload $w --steps 20 --no-cpu-baseline 2>&1 | grep "^{" >> $O/fin_aux.log; done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -o $O/fin_step python bench.py --ncu-step > $O/fin_ncu.log 2>&1
ncu -i $O/fin_step.ncu-rep --page raw --csv > $O/fin_step_raw.csv
rm -f $O/fin_step.ncu-rep
python bench.py > $O/fin_bench.log 2>&1
python bench.py --impl reference --steps 5 --warmup 1 > $O/fin_ref.log 2>&1
tail -c 600 $O/fin_tests.log; tail -2 $O/fin_smoke.log; tail -3 $O/fin_sanitizer.log; wc -l $O/fin_aux.log $O/fin_step_raw.csv; tail -c 300 $O/fin_bench.log; tail -c 200 $O/fin_ref.log
