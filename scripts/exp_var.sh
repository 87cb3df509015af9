#!/bin/bash
# exp_var.sh <tag> <test-variant> <variants...>: parity subset on one variant build, then quick_perf on each (twice, interleaved)
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
tag=$1; tv=$2; shift 2
V=qwen3_tts_cuda_graphs_b200/variants
if [ "$tv" != "-" ]; then
FQ3_LIB_PATH=$V/libfq3_$tv.so timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -5 | tee "gpurun_out/tests_${tag}.log"
fi
rm -f gpurun_out/perf_${tag}.log
for rep in 1 2; do for v in "$@"; do
  echo "variant=$v" | tee -a gpurun_out/perf_${tag}.log
  FQ3_LIB_PATH=$V/libfq3_$v.so timeout 300 python scripts/quick_perf.py 0.6B-Base 64 2>&1 | tail -4 | head -3 | tee -a gpurun_out/perf_${tag}.log
done; done
