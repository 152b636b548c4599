#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_case10_gpu.py tests/test_dns_gpu.py -x -q -m gpu -k "reference_log or restart" 2>&1 | tail -8
timeout 400 python bench.py > gpurun_out/bench_r01_f.json 2> gpurun_out/bench_r01_f.err; tail -c 600 gpurun_out/bench_r01_f.json | head -c 400; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_bench_r01c.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_c.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:splitz_kernel --launch-skip 12 --launch-count 4 -f -o gpurun_out/prof_splitz_r01 python bench.py --steps 1 --warmup 3 --no-cpu --tune split_emulate=2 > gpurun_out/ncu_splitz.log 2>&1; tail -2 gpurun_out/ncu_splitz.log
