#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu --tb=short -x -k "prefill" 2>&1 | tail -25 | tee gpurun_out/tests_dense.log
timeout 300 python scripts/prefill_perf.py 0.6B-Base 2>&1 | tail -4 | tee gpurun_out/prefill_perf.log
