#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000

for c in 0 147; do
  echo "=== CTA $c"
  FQ3_PROF=$c timeout 200 python scripts/phase_prof.py 2>&1 | tail -34
done | tee gpurun_out/phase_prof_ring96.log
