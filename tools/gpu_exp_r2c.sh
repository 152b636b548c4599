#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dns_gpu.py tests/test_march_gpu.py tests/test_case10_gpu.py tests/test_abi.py -x -q -m gpu 2>&1 | tail -6
for t in lazy_scale=0 lazy_scale=1 march_pf=1 march_pf=1,march_red=1; do
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --tune $t > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -c 300 gpurun_out/r2_bench_d.err
  python -c "
import json;d=json.load(open('gpurun_out/r2_bench_d.json'));print('$t', round(d['value'],3),round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items()})"
done
