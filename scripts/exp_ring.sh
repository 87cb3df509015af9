#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
for r in 64 96 128 144 160 176; do
  echo "=== FQ3_RING_KB=$r"
  FQ3_RING_KB=$r timeout 300 python scripts/quick_perf.py 0.6B-Base 32 2>&1 | grep "talker step\|frames="
done | tee gpurun_out/perf_ring2.log
