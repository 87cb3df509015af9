#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
rm -f gpurun_out/ring_1.7B.log
for kb in 0 96 112 128 144; do
  echo "FQ3_RING_KB=$kb" | tee -a gpurun_out/ring_1.7B.log
  if [ $kb = 0 ]; then unset FQ3_RING_KB; else export FQ3_RING_KB=$kb; fi
  timeout 300 python scripts/quick_perf.py 1.7B-Base 32 2>&1 | tail -4 | head -3 | tee -a gpurun_out/ring_1.7B.log
done
