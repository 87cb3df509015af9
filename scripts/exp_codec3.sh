#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_codec_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -8 | tee gpurun_out/tests_codec_tc5b.log
for v in 1 0; do FQ3C_TCGEN05=$v timeout 300 python scripts/codec_ops.py 33 2>&1 | tail -24 | head -24; done | tee gpurun_out/codec_ops2.log
