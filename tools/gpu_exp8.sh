#!/bin/bash
# 8 GPUs: parity of the split-z operators on thin slabs (96 planes), then the C3 and C4 substeps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TLAB_TUNE="splitz=1" TLAB_SHAPE=32,32,768 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/dist_gpu_worker.py 2>&1 | grep -E "DIST_|Error|error|assert" | head -10
run() {
  name=$1; shift
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --no-cpu "$@" > gpurun_out/bench8_$name.json 2> gpurun_out/bench8_$name.err
  python - $name <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open('gpurun_out/bench8_%s.json'%n) if l.startswith('{')][-1])
    print(n, round(d['ms_per_step'],2), round(d['value'],2), d['config'].get('z_operators','')[:20], {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items()}, 'e2e', d['e2e'].get('value'))
except Exception as e:
    print(n, 'failed', e); print(open('gpurun_out/bench8_%s.err'%n).read()[-1500:])
P
}
run c3_splitz1 --steps 5 --warmup 3
run c4_splitz1 --steps 3 --warmup 3 --nx 2048 --ny 1024 --nz 2048
