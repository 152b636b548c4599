#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_march_gpu.py tests/test_operators_gpu.py tests/test_dns_gpu.py tests/test_benchsize_gpu.py -x -q -m gpu > gpurun_out/unsc_tests.log 2>&1
tail -4 gpurun_out/unsc_tests.log
timeout 300 python tests/err_levels.py 2>&1 | grep "^dir"
for t in "march=1" "march=2"; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --tune $t > gpurun_out/unsc_bench_$t.json 2> gpurun_out/unsc_bench_$t.err
  python - "$t" <<'P'
import json,sys
t=sys.argv[1]
try:
    s=open('gpurun_out/unsc_bench_%s.json'%t).read(); d=json.loads(s[s.index('{"metric'):].splitlines()[0])
    print(t, round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items()})
except Exception as e:
    print(t, 'failed', e); print(open('gpurun_out/unsc_bench_%s.err'%t).read()[-1500:])
P
done
