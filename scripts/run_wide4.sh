export FQ3_WATCHDOG_MS=3000
python scripts/wide_debug.py 0.6B-Base 28 5 16 14,60,137,201 2>&1 | tail -1
python scripts/wide_debug.py 0.6B-Base 2 2 7 250,60,97,49,48,145,96 2>&1 | tail -1
FQ3_PROF=77 python scripts/wide_prof.py 16 300 2>&1 | tail -13
python scripts/quick_perf.py 0.6B-Base 64 14 2>&1 | tail -1
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_parity_gpu.py -q -m gpu -x --tb=short 2>&1 | tail -4
