#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
FQ3_LIB_PATH=qwen3_tts_cuda_graphs_b200/variants/libfq3_checks.so timeout 1200 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -8 | tee gpurun_out/tests_checks_build.log
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -3
