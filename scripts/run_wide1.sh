export FQ3_WATCHDOG_MS=3000
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -x --tb=short -k "wide or two_streams or normal_program" 2>&1 | tail -30 > gpurun_out/wide_tests1.log; tail -5 gpurun_out/wide_tests1.log
timeout 600 python scripts/batch_perf.py 0.6B-Base 16 1,4,8,16 > gpurun_out/wide_perf1.log 2>&1; cat gpurun_out/wide_perf1.log
