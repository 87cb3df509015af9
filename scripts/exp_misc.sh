#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=5000
timeout 300 python scripts/batch_perf.py 0.6B-Base 32 2>&1 | tail -4 | tee gpurun_out/batch_perf.log
timeout 400 python scripts/quick_perf.py 1.7B-Base 32 2>&1 | tail -5 | tee gpurun_out/perf_1.7B.log
timeout 400 python scripts/batch_perf.py 1.7B-Base 16 2>&1 | tail -4 | tee -a gpurun_out/batch_perf.log
