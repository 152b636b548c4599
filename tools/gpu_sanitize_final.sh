#!/bin/bash
# compute-sanitizer over the kernels that are new in the final tree: circulant form (x whole-line, z marching), peeled marching steps in y,
# split-z kernels in circulant form on two virtual slabs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_step.py circ_peel split_circ > gpurun_out/sanitize_memcheck_final.log 2>&1
echo "== memcheck rc=$?"; grep -E "variant|SANITIZE_STEP_OK|ERROR SUMMARY|Error|error" gpurun_out/sanitize_memcheck_final.log | head -12
timeout 150 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_step.py circ_peel > gpurun_out/sanitize_racecheck_final.log 2>&1
echo "== racecheck rc=$?"; grep -E "variant|SANITIZE_STEP_OK|RACECHECK SUMMARY|Error|error" gpurun_out/sanitize_racecheck_final.log | head -12
