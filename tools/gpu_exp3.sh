#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_poisson_gpu.py tests/test_dns_gpu.py -x -q -m gpu -k "not tma and not virtual" 2>&1 | tail -5
timeout 200 python tools/bench_ops.py --poisson --no-lines --shape 128,512,1024 --iters 10 --warmup 3 --json gpurun_out/poisson_p8like.json 2>&1 | grep -v "^$" | tail -16
timeout 300 python tools/bench_ops.py --poisson --no-lines --shape 1024,512,1024 --iters 5 --warmup 2 --json gpurun_out/poisson_c3.json 2>&1 | tail -16
