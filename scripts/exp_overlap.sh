#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_engine_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -3 | tee gpurun_out/tests_overlap.log
FQ3_GRID=128 FQ3_OVERLAP_CODEC=1 timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -3 | tee -a gpurun_out/tests_overlap.log
rm -f gpurun_out/overlap_bench.log
for mode in "0 148" "1 128" "0 148" "1 128"; do set -- $mode
  echo "FQ3_OVERLAP_CODEC=$1 FQ3_GRID=$2" | tee -a gpurun_out/overlap_bench.log
  FQ3_OVERLAP_CODEC=$1 FQ3_GRID=$2 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batched-streams 0 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ttfa', round(d['ttfa_ms']['mean'],2), 'ms/frame', round(d['decode_ms_per_frame'],3))" | tee -a gpurun_out/overlap_bench.log
done
