#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dns_gpu.py -x -q -m gpu -k "virtual_slabs" 2>&1 | tail -6
for t in split_emulate=8,march=0 split_emulate=8,march=1 split_emulate=2,march=0 split_emulate=2,march=1; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --tune $t > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -c 300 gpurun_out/r2_bench_d.err
  python -c "
import json;d=json.load(open('gpurun_out/r2_bench_d.json'));print('$t', round(d['value'],3),round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items()})"
done
