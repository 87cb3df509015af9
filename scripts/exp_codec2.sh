#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do FQ3C_TCGEN05=$v timeout 300 python scripts/codec_ops.py 33 2>&1 | tail -24; done | tee gpurun_out/codec_ops.log
