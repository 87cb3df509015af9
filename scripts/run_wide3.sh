export FQ3_WATCHDOG_MS=3000
python scripts/quick_perf.py 0.6B-Base 64 14 2>&1 | tail -1
FQ3_PROF=0 python scripts/wide_prof.py 16 14 2>&1 | tail -9
FQ3_PROF=77 python scripts/wide_prof.py 16 300 2>&1 | tail -9
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_parity_gpu.py -q -m gpu -x --tb=short 2>&1 | tail -8 > gpurun_out/wide_tests3.log; tail -4 gpurun_out/wide_tests3.log
python scripts/wide_debug.py 0.6B-Base 28 5 16 14,60,37,101 2>&1 | tail -1
