#!/bin/bash
# compute-sanitizer over every kernel family on tiny shapes; one-line verdicts go to gpurun_out/sanitizer_summary.txt
# usage (GPU box): bash scripts/sanitize.sh [memcheck racecheck synccheck initcheck]
set -u
mkdir -p gpurun_out
TOOLS="${*:-memcheck racecheck synccheck}"
: > gpurun_out/sanitizer_summary.txt
for tool in $TOOLS; do
  for shape in tiny tiny-moe; do
    log=gpurun_out/sanitizer_${tool}_${shape}.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_target.py $shape > $log 2>&1
    rc=$?
    echo "$tool $shape: exit $rc; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|hazard' $log | tail -2 | tr '\n' ' ')" | tee -a gpurun_out/sanitizer_summary.txt
  done
done
