#!/bin/bash
# poll back-off A/B (FQ3_DEBUG = tries << 16 | ns) + L2 homing microbenchmark + per-warp timeline of the current kernel
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
timeout 120 ./bench_micro/l2_home > gpurun_out/l2_home.log 2>&1
rm -f gpurun_out/backoff.log
for d in 262176 65568 1048608 262164 262208 6553600; do
  echo "FQ3_DEBUG=$d (tries=$((d>>16)) ns=$((d&65535)))" | tee -a gpurun_out/backoff.log
  FQ3_DEBUG=$d timeout 300 python scripts/quick_perf.py 0.6B-Base 64 2>&1 | tail -2 | head -1 | tee -a gpurun_out/backoff.log
done
FQ3_PROF=-3 timeout 200 python scripts/warp_skew.py > gpurun_out/warp_skew_v8.log 2>&1
