#!/bin/bash
tag=${1:-t}
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -25 | tee "gpurun_out/tests_${tag}.log"
timeout 300 python scripts/quick_perf.py 0.6B-Base 64 2>&1 | tail -5 | tee gpurun_out/perf_${tag}.log
FQ3_PROF=0 timeout 200 python scripts/pred_prof.py 2>&1 | tail -9 | tee gpurun_out/pred_prof_${tag}.log
