#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_codec_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -3 | tee gpurun_out/tests_2cta.log
FQ3C_TC5_MINK=64 timeout 900 python -m pytest tests/test_codec_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -3 | tee -a gpurun_out/tests_2cta.log
for v in 256 64 256 64; do
echo "FQ3C_TC5_MINK=$v" | tee -a gpurun_out/twocta_perf.log
FQ3C_TC5_MINK=$v timeout 300 python scripts/codec_time.py 2>&1 | tail -2 | tee -a gpurun_out/twocta_perf.log
done
FQ3C_TC5_MINK=64 timeout 300 python scripts/codec_ops.py 2>&1 | tail -24 | head -12 | tee gpurun_out/codec_ops4.log
