#!/bin/bash
# one GPU call: parity of the TMA kernels, then the substep breakdown for a few tilings, then an ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dns_gpu.py -x -q -m gpu -k "tma" 2>&1 | tail -5
for t in "tma=1,lines_yz=8" "tma=1,lines_yz=4" "tma=1,lines_yz=4,tma_l2=1" "tma=1,lines_yz=4,tma_l2=2" "tma=1,lines_yz=8,tma_l2=2"; do
  timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu --tune $t > gpurun_out/bench_tma_$t.json 2> gpurun_out/bench_tma_$t.err
  python - "$t" <<'P'
import json,sys
t=sys.argv[1]
try:
    d=json.load(open('gpurun_out/bench_tma_%s.json'%t))
    print(t, round(d['ms_per_step'],1), {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items() if 'y' in k or 'z' in k})
except Exception as e:
    print(t, 'failed', e)
P
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lines2_strided_tma --launch-count 11 -f -o gpurun_out/prof_tma_v1 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_tma_v1.log 2>&1
tail -3 gpurun_out/ncu_tma_v1.log
