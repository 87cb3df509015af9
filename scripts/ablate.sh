#!/bin/bash
# timing ablations of the talker step / predictor (results are WRONG under FQ3_DEBUG != 0; timing only)
for d in 0 1 2 4 8 3 6 7 15; do
  echo "FQ3_DEBUG=$d"; FQ3_DEBUG=$d timeout 200 python scripts/quick_perf.py 0.6B-Base 16 2>&1 | grep -E "talker step|predictor:|frames="
done
