#!/bin/bash
# CompactDirect6 slice + a proper launch list of the timed substeps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_operators_gpu.py tests/test_host_cpu.py -x -q -k "direct or partial or burgers" > gpurun_out/r2_direct_tests.log 2>&1
tail -5 gpurun_out/r2_direct_tests.log
bash tools/gpu_launches_r2.sh > gpurun_out/r2_launches.log 2>&1
tail -5 gpurun_out/r2_launches.log
