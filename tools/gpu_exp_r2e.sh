#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_poisson_gpu.py tests/test_benchsize_gpu.py tests/test_case10_gpu.py tests/test_dns_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== per-GPU share of the Poisson y stage at 8 GPUs (65 x 1024 modes x 512 rows), incl. the singular modes"
POISSON_CFG=3:0 POISSON_IL=0 timeout 300 python tools/bench_ops.py --shape 128,512,1024 --poisson --no-lines --iters 10 --warmup 3 2>&1 | grep -E "OPR_Poisson|y solves"
echo "== C4 share: 256 x 1024 x 2048 (129 x 2048 modes x 1024 rows)"
POISSON_CFG=3:0 POISSON_IL=0 timeout 300 python tools/bench_ops.py --shape 256,1024,2048 --poisson --no-lines --iters 5 --warmup 2 2>&1 | grep -E "OPR_Poisson|y solves"
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -c 300 gpurun_out/r2_bench_d.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_d.json'));print(round(d['value'],3),round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items()})"
