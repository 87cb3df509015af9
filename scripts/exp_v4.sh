#!/bin/bash
# v4 kernel bring-up: parity tests group by group, then quick perf + phase profile
tag=${1:-v4a}
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
for grp in "test_linear" "test_talker_step or test_prefill" "test_predictor or test_frame_loop" "test_streaming or test_min_new or test_static or test_two_streams"; do
  name=$(echo "$grp" | tr ' ' '_')
  echo "=== $grp"
  timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "$grp" --tb=short 2>&1 | tail -40 | tee "gpurun_out/tests_${tag}_${name}.log"
done
timeout 300 python scripts/quick_perf.py 0.6B-Base 64 2>&1 | tail -6 | tee gpurun_out/perf_${tag}.log
FQ3_PROF=0 timeout 200 python scripts/phase_prof.py 2>&1 | tail -40 | tee gpurun_out/phase_prof_${tag}.log
