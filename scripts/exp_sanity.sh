#!/bin/bash
# memcheck of the decode kernel on the tiny config (bounded), then 1.7B and batched timings
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=120000
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_engine_gpu.py -q -m gpu -x --tb=short \
  -k "tiny and (talker_step_matches or prefill_matches or predictor or two_streams or long_context)" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/memcheck.log
tail -6 gpurun_out/memcheck.log
export FQ3_WATCHDOG_MS=3000
timeout 300 python scripts/quick_perf.py 1.7B-Base 32 2>&1 | tail -5 | tee gpurun_out/perf_1.7B_r01d.log
timeout 300 python scripts/batch_perf.py 0.6B-Base 48 2>&1 | tail -3 | tee gpurun_out/batch_r01d.log
timeout 300 python scripts/batch_perf.py 1.7B-Base 32 2>&1 | tail -3 | tee -a gpurun_out/batch_r01d.log
