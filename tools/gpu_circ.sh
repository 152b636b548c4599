#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_operators_gpu.py tests/test_march_gpu.py tests/test_dns_gpu.py -x -q -m gpu > gpurun_out/circ_tests.log 2>&1
tail -4 gpurun_out/circ_tests.log
for t in circ=1 circ=0; do
timeout 300 python tools/bench_ops.py --shape 1024,512,512 --only "_X " --persist 0 --iters 10 --warmup 3 --tune $t 2>&1 | grep OPR
done
timeout 300 python tools/bench_ops.py --shape 1024,512,512 --only "_Z " --persist 0 --iters 10 --warmup 3 --tune circ=1,march=0 2>&1 | grep OPR
timeout 300 python tools/bench_ops.py --shape 1024,512,512 --only "_Z " --persist 0 --iters 10 --warmup 3 --tune circ=0,march=0 2>&1 | grep OPR
timeout 300 python tools/bench_ops.py --shape 1024,512,512 --only "_Z " --persist 0 --iters 10 --warmup 3 --tune march=1 2>&1 | grep OPR
