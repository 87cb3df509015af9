#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short -x -k "prefill or dense or model or stream or clone" 2>&1 | tail -4 | tee gpurun_out/tests_pfgraph.log
timeout 300 python scripts/prefill_perf.py 0.6B-Base 2>&1 | tail -4 | tee gpurun_out/prefill_perf3.log
FQ3C_GRAPH=0 timeout 300 python scripts/prefill_perf.py 0.6B-Base 2>&1 | tail -4 | tee -a gpurun_out/prefill_perf3.log
timeout 300 python scripts/ttfa_breakdown.py 2>&1 | tail -6 | tee gpurun_out/ttfa_r01d.log
