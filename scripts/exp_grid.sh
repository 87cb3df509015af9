#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
rm -f gpurun_out/grid_sweep.log
for g in 148 144 136 128 112 96; do
  echo "FQ3_GRID=$g" | tee -a gpurun_out/grid_sweep.log
  FQ3_GRID=$g timeout 300 python scripts/quick_perf.py 0.6B-Base 32 2>&1 | tail -4 | head -3 | tee -a gpurun_out/grid_sweep.log
done
