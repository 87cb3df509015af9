#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=20000
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "test_predictor and tiny" -x --tb=line 2>&1 | grep -v "^$" | head -60 > gpurun_out/sanitizer_v3b.log
tail -40 gpurun_out/sanitizer_v3b.log
export FQ3_WATCHDOG_MS=3000
timeout 300 python scripts/quick_perf.py 0.6B-Base 64 2>&1 | tail -6 | tee gpurun_out/perf_v3b.log
FQ3_PROF=0 timeout 200 python scripts/phase_prof.py 2>&1 | tail -14 | tee gpurun_out/phase_prof_v3b.log
