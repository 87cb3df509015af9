#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 120 ./bench_micro/exch_lat 2>&1 | tee gpurun_out/exch_lat_r1.log
for st in 6 9 12; do
  echo "FQ3_STAGES=$st"; FQ3_STAGES=$st timeout 200 python scripts/quick_perf.py 0.6B-Base 16 2>&1 | grep -E "talker step|predictor:|frames=|prefill"
done | tee gpurun_out/stages_sweep.log
FQ3_STAGES=12 FQ3_PROF=0 timeout 200 python scripts/phase_prof.py 2>&1 | tail -12 | tee gpurun_out/phase_prof_st12.log
