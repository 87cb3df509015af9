export FQ3_WATCHDOG_MS=3000
python scripts/wide_debug.py 0.6B-Base 28 5 16 14,60,37,101 2>&1 | tail -2
python scripts/quick_perf.py 0.6B-Base 64 14 2>&1 | tail -1
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_parity_gpu.py -q -m gpu -x --tb=short 2>&1 | tail -8 > gpurun_out/wide_tests2.log; tail -4 gpurun_out/wide_tests2.log
timeout 600 python scripts/batch_perf.py 0.6B-Base 16 1,4,8,16,32,64 > gpurun_out/wide_perf2.log 2>&1; cat gpurun_out/wide_perf2.log | cut -c1-150
