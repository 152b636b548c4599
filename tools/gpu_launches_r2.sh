#!/bin/bash
# launch list of the substeps: everything except torch's element-wise kernels (field synthesis) and the one-off
# initialisation kernels of the Poisson solver (ncu matches the regex against the function base name)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k 'regex:^(?!(vectorized_elementwise_kernel|elementwise_kernel|unrolled_elementwise_kernel|poisson_fundamental_kernel|poisson_relayout_kernel|index_elementwise_kernel|reduce_kernel)).*$' -c 400 --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-parity > gpurun_out/bench_under_ncu_r02.log 2>&1
tail -2 gpurun_out/bench_under_ncu_r02.log | cut -c1-200
wc -l gpurun_out/launches_bench_r02.csv
