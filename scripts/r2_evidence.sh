#!/bin/bash
# round-2 evidence: TF32 peak, racecheck / synccheck over the op tests of the tcgen05 kernels, ncu launch list of a step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/measure_tf32_peak.py > gpurun_out/r2_tf32_peak.json 2>/dev/null; cat gpurun_out/r2_tf32_peak.json
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_ops.py -x -q -k "dwpw and (16-16-96 or 26-26-288 or 61-45) or se_parity or fused_squeeze or (pw_parity and (3-3-16-16 or 3-2-13-13-512 or 3-8-40-40 or 4-3-16-16 or 4-2-13-13-512 or 4-8-40-40 or 4-8-52-52))" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; tail -5 gpurun_out/r2_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_ops.py -x -q -k "dwpw and (16-16-96 or 26-26-288 or 61-45) or se_parity or fused_squeeze" > gpurun_out/r2_sanitizer_synccheck.log 2>&1; tail -4 gpurun_out/r2_sanitizer_synccheck.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py tests/test_gpu_loss.py tests/test_gpu_train.py -x -q -k "dwpw or se_parity or fused_squeeze or fused_loss or sparse or adam" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/r2_sanitizer_memcheck.log
