#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_march_gpu.py tests/test_operators_gpu.py tests/test_dns_gpu.py tests/test_benchsize_gpu.py -x -q -m gpu > gpurun_out/circ3_tests.log 2>&1
tail -4 gpurun_out/circ3_tests.log
for t in "circ=1" ; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --tune $t > gpurun_out/circ3_bench_$t.json 2> gpurun_out/circ3_bench_$t.err
  python - "$t" <<'P'
import json,sys
t=sys.argv[1]
try:
    s=open('gpurun_out/circ3_bench_%s.json'%t).read(); d=json.loads(s[s.index('{"metric'):].splitlines()[0])
    print(t, round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items()})
    print(json.dumps(d['config'].get('parity') or d.get('parity'))[:600])
except Exception as e:
    print(t, 'failed', e); print(open('gpurun_out/circ3_bench_%s.err'%t).read()[-1500:])
P
done
