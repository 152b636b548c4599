#!/bin/bash
# compute-sanitizer over small RK steps through every kernel family: memcheck, then racecheck (shared-memory
# hazards: exchange areas of the line kernels, stash and rings of the marching kernels, team scans of the Poisson y solves)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_step.py "$@" > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "variant|SANITIZE_STEP_OK|ERROR SUMMARY|RACECHECK SUMMARY|Error|error" gpurun_out/sanitize_$tool.log | head -30
done
