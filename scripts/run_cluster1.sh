export FQ3_WATCHDOG_MS=3000
for c in 8 4 2 1; do
  echo "== FQ3_WIDE_CLUSTER=$c"
  FQ3_WIDE_CLUSTER=$c timeout 300 python scripts/wide_debug.py 0.6B-Base 2 2 16,5 14,60,137,201 2>&1 | tail -2
  FQ3_WIDE_CLUSTER=$c FQ3_PROF=0 timeout 300 python scripts/wide_prof.py 16 14 2>&1 | grep -v "leader\|multi" | tail -9
done
