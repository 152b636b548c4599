#!/bin/bash
# bench line at N GPUs exactly as the driver launches it (parity record, C4 block at N = 8, e2e included)
N=$1; TAG=$2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu_c3_${TAG}.json 2> gpurun_out/bench_${N}gpu_c3_${TAG}.err
python - $N $TAG <<'P'
import json,sys
n,tag=sys.argv[1],sys.argv[2]
try:
    s=open('gpurun_out/bench_%sgpu_c3_%s.json'%(n,tag)).read(); d=json.loads(s[s.index('{"metric'):].splitlines()[0])
    print(n, round(d['ms_per_step'],2), round(d['value'],2), {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items()}, 'e2e', d['e2e'].get('value'))
    print('parity', json.dumps(d.get('parity'))[:700])
    for k in d:
        if 'c4' in k.lower(): print(k, json.dumps(d[k])[:900])
except Exception as e:
    print(n, 'failed', e); print(open('gpurun_out/bench_%sgpu_c3_%s.err'%(n,tag)).read()[-2500:])
P
