#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
for d in $((4<<16|32)) $((0<<16|100)) $((0<<16|300)) $((0<<16|1000)) $((100<<16|0)); do
  echo "=== FQ3_DEBUG=$d"
  FQ3_DEBUG=$d timeout 300 python scripts/quick_perf.py 0.6B-Base 32 2>&1 | tail -4 | head -3
done | tee gpurun_out/perf_sleep.log
