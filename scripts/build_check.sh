#!/bin/bash
# compile libfq3.so and report registers / spills / local-memory instructions of the stream kernel
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared --expt-relaxed-constexpr -diag-suppress 550 -Xptxas -v -o qwen3_tts_cuda_graphs_b200/libfq3.so qwen3_tts_cuda_graphs_b200/csrc/fq3_api.cu 2>&1 | grep -v "^$" | grep -A2 "error\|stream_kernel" | grep -v "^--"
rm -rf /tmp/cub && mkdir -p /tmp/cub && (cd /tmp/cub && cuobjdump -xelf all $OLDPWD/qwen3_tts_cuda_graphs_b200/libfq3.so >/dev/null && nvdisasm -g -c *.cubin > /tmp/dis.txt 2>/dev/null)
start=$(grep -n "text._ZN3fq317fq3_stream_kernel" /tmp/dis.txt | head -1 | cut -d: -f1)
awk -v s=$start 'NR>=s' /tmp/dis.txt | awk '/\/\/## File/ {line=$0} /LDL|STL/ {print line}' | sed 's/.*fq3_kernel.cuh", line \([0-9]*\).*/\1/' | sort -n | uniq -c | sort -k2 -n | awk '{printf "%s:%s ", $2, $1} END {print ""}'
