#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
rm -f gpurun_out/l2pf.log
for m in 1.7B-Base 0.6B-Base; do for v in 0 1 0 1; do
  echo "$m FQ3_LLMODE=$v" | tee -a gpurun_out/l2pf.log
  FQ3_LLMODE=$v timeout 300 python scripts/quick_perf.py $m 32 2>&1 | tail -4 | head -3 | tee -a gpurun_out/l2pf.log
done; done
