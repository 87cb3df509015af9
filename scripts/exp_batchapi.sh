#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
timeout 120 ./bench_micro/l2_home > gpurun_out/l2_home.log 2>&1
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -8 | tee gpurun_out/tests_batchapi.log
timeout 300 python scripts/quick_perf.py 0.6B-Base 64 2>&1 | tail -2 | head -1
