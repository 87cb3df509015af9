#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
for r in 64 80 96 112 128 176; do
  echo "=== FQ3_RING_KB=$r"
  FQ3_RING_KB=$r timeout 300 python scripts/quick_perf.py 0.6B-Base 32 2>&1 | tail -4 | head -3
done | tee gpurun_out/perf_ring.log
