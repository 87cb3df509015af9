#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
timeout 600 python -m pytest tests/test_codec_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short 2>&1 | tail -60 | tee gpurun_out/tests_codec_model.log
timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench_first.log
