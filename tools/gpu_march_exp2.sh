#!/bin/bash
# C4 line lengths on one GPU: y lines of 1024 points and x lines of 2048 points, whole-line kernels vs marching panels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_abi.py -x -q -m gpu 2>&1 | tail -4
for m in 0 2; do
echo "== march=$m, 1024 x 1024 x 256"
timeout 300 python tools/bench_ops.py --shape 1024,1024,256 --iters 10 --warmup 3 --persist 0 --only "Partial_Y P1|Burgers_Y|Partial_Z P1|Burgers_Z" --tune march=$m 2>&1 | grep OPR
done
echo "== x lines of 2048: 2048 x 512 x 256"
timeout 300 python tools/bench_ops.py --shape 2048,512,256 --iters 10 --warmup 3 --persist 0 --only "Partial_X P1|Burgers_X" 2>&1 | grep OPR
