#!/bin/bash
# Run the GPU parity tests group by group (a device fault in one group must not poison the next).
export FQ3_WATCHDOG_MS=${FQ3_WATCHDOG_MS:-3000}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for grp in "test_linear" "test_talker_step or test_prefill" "test_predictor or test_frame_loop" "test_streaming or test_min_new or test_static or test_two_streams"; do
  name=$(echo "$grp" | tr ' ' '_')
  echo "=== $grp"
  timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "$grp" -x --tb=short 2>&1 | tail -40 | tee "gpurun_out/tests_${name}.log"
  [ ${PIPESTATUS[0]} -ne 0 ] && rc=1
done
exit $rc
