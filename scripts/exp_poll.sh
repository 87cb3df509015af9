#!/bin/bash
# A/B of the polling flavours (FQ3_LLMODE bit 3 = single-request polling) + a parity subset
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
tag=${1:-t}
timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -5 | tee "gpurun_out/tests_${tag}.log"
for mode in 0 8 0 8; do
  echo "LLMODE=$mode" | tee -a gpurun_out/perf_${tag}.log
  FQ3_LLMODE=$mode timeout 300 python scripts/quick_perf.py 0.6B-Base 64 2>&1 | tail -4 | head -3 | tee -a gpurun_out/perf_${tag}.log
done
