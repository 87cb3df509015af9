#!/bin/bash
# build libfq3.so with extra -D flags into variants/<name>/ (a copy of the package dir with its own .so), for A/B runs:
#   scripts/build_variant.sh pre0 -DFQ3_PRE=0 ; PYTHONPATH=variants/pre0 python scripts/quick_perf.py
name=$1; shift
mkdir -p variants/$name
rm -rf variants/$name/qwen3_tts_cuda_graphs_b200 variants/$name/include
cp -r qwen3_tts_cuda_graphs_b200 include variants/$name/
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared --expt-relaxed-constexpr -diag-suppress 550 -diag-suppress 177 "$@" \
  -o variants/$name/qwen3_tts_cuda_graphs_b200/libfq3.so variants/$name/qwen3_tts_cuda_graphs_b200/csrc/fq3_api.cu && touch variants/$name/qwen3_tts_cuda_graphs_b200/libfq3.so && echo built $name
