#!/bin/bash
# build_variant.sh <name> [-DFLAG ...]: a differently compiled libfq3 under qwen3_tts_cuda_graphs_b200/variants/ (FQ3_LIB_PATH selects it)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p qwen3_tts_cuda_graphs_b200/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared --expt-relaxed-constexpr -diag-suppress 550 \
  -Xptxas -v "$@" -o qwen3_tts_cuda_graphs_b200/variants/libfq3_$name.so qwen3_tts_cuda_graphs_b200/csrc/fq3_api.cu 2>&1 | grep -A2 "fq3_stream_kernelILb0" | grep -E "spill|registers"
