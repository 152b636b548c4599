#!/bin/bash
# per-kernel times of the split-z operators at the 8-GPU slab shape of C3 (1024x512x128 per slab): virtual slabs on one GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k 'regex:splitz' -c 240 --log-file gpurun_out/launches_splitz_emul8.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-parity --tune split_emulate=8 > gpurun_out/bench_emul8_ncu.log 2>&1
tail -c 300 gpurun_out/bench_emul8_ncu.log
python tools/launch_summary.py gpurun_out/launches_splitz_emul8.csv gpurun_out/launches_splitz_emul8_summary.json > /dev/null
python - <<'P'
import json
d=json.load(open('gpurun_out/launches_splitz_emul8_summary.json'))
for k in d['kernels']: print(k['kernel'][:90], k['launches'], round(k['ms']/k['launches'],4))
P
