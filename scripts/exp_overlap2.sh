#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
timeout 1200 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -3 | tee gpurun_out/tests_overlap2.log
FQ3_OVERLAP_CODEC=0 timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -2 | tee -a gpurun_out/tests_overlap2.log
rm -f gpurun_out/overlap_bench2.log
for v in 1 0 1 0; do
  echo "FQ3_OVERLAP_CODEC=$v" | tee -a gpurun_out/overlap_bench2.log
  FQ3_OVERLAP_CODEC=$v timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batched-streams 0 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ttfa', round(d['ttfa_ms']['mean'],2), 'ms/frame', round(d['decode_ms_per_frame'],3), 'frac', round(d['roofline']['frac'],4))" | tee -a gpurun_out/overlap_bench2.log
done
timeout 300 python scripts/quick_perf.py 0.6B-Base 32 2>&1 | tail -2 | head -1 | tee -a gpurun_out/overlap_bench2.log
