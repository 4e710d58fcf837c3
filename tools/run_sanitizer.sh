#!/bin/bash
# compute-sanitizer over the hot path (GPU box).  Writes one summary per tool into gpurun_out/sanitizer/.
#   tools/run_sanitizer.sh [cases...]     default: price crossing greeks moments xgpu
set -u
OUT=gpurun_out/sanitizer
mkdir -p $OUT
CASES=${@:-price crossing greeks moments xgpu}
SAN=${SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
for tool in memcheck racecheck synccheck initcheck; do
  for c in $CASES; do
    log=$OUT/${tool}_${c}.log
    timeout 900 $SAN --tool $tool --print-limit 20 python tools/sanitizer_cases.py $c > $log 2>&1
    rc=$?
    summary=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $log | tail -1)
    ok=$(grep -c "sanitizer cases ok" $log)
    echo "$tool $c exit=$rc cases_ok=$ok :: ${summary:-no summary line}" | tee -a $OUT/summary.txt
  done
done
