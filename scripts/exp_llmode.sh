#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
for m in 0 40 56 64 $((64+40)) $((56+5)); do
  echo "=== FQ3_LLMODE=$m"
  FQ3_LLMODE=$m timeout 300 python scripts/quick_perf.py 0.6B-Base 16 2>&1 | grep "talker step\|predictor\|frames="
done | tee gpurun_out/perf_llmode2.log
