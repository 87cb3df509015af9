#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -4 | tee gpurun_out/tests_combine.log
FQ3_LIB_PATH=qwen3_tts_cuda_graphs_b200/variants/libfq3_checks.so timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -4 | tee gpurun_out/tests_checks_build.log
timeout 400 python scripts/frame_cost.py 2>&1 | tail -6 | tee gpurun_out/frame_cost_v7j.log
