#!/bin/bash
# 2 GPUs: what the NVLink stores of the split-z phase 1 cost (trimmed / untrimmed exchange, stores kept local)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in "split_trim=0" "split_trim=1" "split_trim=1,split_local=1" $EXTRA; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-extra --no-parity --no-c4 --tune $t > gpurun_out/ab2_$t.json 2> gpurun_out/ab2_$t.err
  python - "$t" <<'P'
import json,sys
t=sys.argv[1]
try:
    s=open('gpurun_out/ab2_%s.json'%t).read(); d=json.loads(s[s.index('{"metric'):].splitlines()[0])
    print(t, round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items()})
except Exception as e:
    print(t, 'failed', e); print(open('gpurun_out/ab2_%s.err'%t).read()[-1500:])
P
done
