#!/bin/bash
# compute-sanitizer over a small run of every kernel (tools/sanitize_target.py), on the GPU box:
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Writes gpurun_out/sanitize_<tool>_<part>.log; the last lines carry the error summary.
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for part in robot host natives crowd; do
    log=gpurun_out/sanitize_${tool}_${part}.log
    timeout 400 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py $part > $log 2>&1
    echo "== $tool $part rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok' $log | tr '\n' ' ')"
  done
done
