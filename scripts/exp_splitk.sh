#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
timeout 900 python -m pytest tests/test_codec_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -5 | tee gpurun_out/tests_splitk.log
timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu --tb=short -x -k "prefill or dense" 2>&1 | tail -3 | tee -a gpurun_out/tests_splitk.log
for sk in 1 0; do
echo "FQ3C_SPLITK=$sk" | tee -a gpurun_out/splitk_perf.log
FQ3C_SPLITK=$sk timeout 300 python scripts/prefill_perf.py 0.6B-Base 2>&1 | tail -4 | tee -a gpurun_out/splitk_perf.log
FQ3C_SPLITK=$sk timeout 300 python scripts/codec_time.py 2>&1 | tail -3 | tee -a gpurun_out/splitk_perf.log
done
