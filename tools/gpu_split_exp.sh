#!/bin/bash
# split-z operators: virtual slabs on one GPU, then 2 GPUs (parity + substep time with and without)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dns_gpu.py -x -q -m gpu -k "virtual_slabs" 2>&1 | tail -15 || exit 1
timeout 400 python -m pytest tests/test_dist_gpu.py -x -q -m gpu -k "split_z" 2>&1 | tail -15
for t in "splitz=1" "splitz=0"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --tune $t > gpurun_out/bench2_$t.json 2> gpurun_out/bench2_$t.err
  python - "$t" <<'P'
import json,sys
t=sys.argv[1]
try:
    d=json.loads([l for l in open('gpurun_out/bench2_%s.json'%t) if l.startswith('{')][-1])
    print(t, round(d['ms_per_step'],1), round(d['value'],2), {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items()})
except Exception as e:
    print(t, 'failed', e); print(open('gpurun_out/bench2_%s.err'%t).read()[-1500:])
P
done
