#!/bin/bash
# compute-sanitizer over small solves (one GPU): memcheck and racecheck, default kernels and the opt-in TMA projection.
mkdir -p gpurun_out
rm -f gpurun_out/sanitizer.log
for tool in memcheck racecheck; do
  for tma in 0 1; do
    echo "== compute-sanitizer --tool $tool   J2P_PROJ_TMA=$tma" >> gpurun_out/sanitizer.log
    J2P_PROJ_TMA=$tma timeout 400 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_child.py >> gpurun_out/sanitizer.log 2>&1
    echo "exit code $?" >> gpurun_out/sanitizer.log
  done
done
grep -E "^==|exit code|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard" gpurun_out/sanitizer.log | head -40
