#!/bin/bash
# single-GPU C3 substep for a few tilings / schedules of the line kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in "$@"; do
  timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu --tune $t > gpurun_out/bench_sweep_$t.json 2> gpurun_out/bench_sweep_$t.err
  python - "$t" <<'P'
import json,sys
t=sys.argv[1]
try:
    d=json.load(open('gpurun_out/bench_sweep_%s.json'%t))
    print(t, round(d['ms_per_step'],1), {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items() if 'burgers' in k or 'partial' in k})
except Exception as e:
    print(t, 'failed', e)
P
done
