#!/bin/bash
# round 2, final tree: the whole GPU suite, then the bench line and the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_gpu_suite_final.log
timeout 500 python bench.py > gpurun_out/bench_1gpu_c3_r02.json 2> gpurun_out/bench_1gpu_c3_r02.err; tail -c 300 gpurun_out/bench_1gpu_c3_r02.err
python -c "
import json;d=json.load(open('gpurun_out/bench_1gpu_c3_r02.json'));print(round(d['value'],3),round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items()}); print(d['e2e']); print({k:round(v.get('frac_fused_bytes',v.get('frac',0)),3) for k,v in d['roofline']['per_class'].items()})"
