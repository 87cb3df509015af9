export FQ3_WATCHDOG_MS=3000
python scripts/wide_debug.py 0.6B-Base 28 5 16 14,60,137,201 2>&1 | tail -1
python scripts/wide_debug.py 0.6B-Base 2 2 7,1 250,60,97,49,48,145,96 2>&1 | tail -2
timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -x --tb=short -k "wide or normal_program or two_streams" 2>&1 | tail -2
for v in base ""; do echo "variant=$v"; FQ3_VARIANT=$v python scripts/batch_perf.py 0.6B-Base 16 8,16 2>&1 | tail -2 | cut -c1-100; done
MAX_SEQ=1024 python scripts/batch_perf.py 0.6B-Base 16 16 2>&1 | tail -1 | cut -c1-100
