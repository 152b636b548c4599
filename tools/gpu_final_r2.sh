#!/bin/bash
# round 2: the bench line, the reference arm, and the ncu evidence of the same command (launch list, DRAM bytes per launch,
# --set full of the dominant kernels, summarised on the box: the reports themselves are too large to bring back)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$1" = "suite" ]; then timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_gpu_suite.log; fi
timeout 500 python bench.py > gpurun_out/bench_1gpu_c3_r02.json 2> gpurun_out/bench_1gpu_c3_r02.err; tail -c 300 gpurun_out/bench_1gpu_c3_r02.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; tail -c 300 gpurun_out/bench_ref_r02.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_under_ncu_r02.log 2>&1
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv -k "regex:lines2_|poisson_team" --launch-skip 66 --launch-count 22 --log-file gpurun_out/dram_bench_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_under_ncu_dram_r02.log 2>&1
# one substep = 22 line-kernel / Poisson launches; --set full on the Burgers and derivative kernels of each direction and the y solves
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:lines2_march|lines2_contig|lines2_strided|poisson_team" --launch-skip 66 --launch-count 22 -f -o /tmp/prof_bench_r02 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_r02.log 2>&1; tail -2 gpurun_out/ncu_full_r02.log
python tools/ncu_key.py /tmp/prof_bench_r02.ncu-rep > gpurun_out/ncu_full_bench_r02.txt 2>&1
ncu -i /tmp/prof_bench_r02.ncu-rep --page raw --csv > gpurun_out/ncu_full_bench_r02_raw.csv 2>/dev/null
ls -la /tmp/prof_bench_r02.ncu-rep gpurun_out/ | head -20
python -c "
import json;d=json.load(open('gpurun_out/bench_1gpu_c3_r02.json'));print(round(d['value'],3),round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items()}); print(d['cpu_baseline']); print(d['e2e'])
r=json.load(open('gpurun_out/bench_ref_r02.json')); print('reference arm', r['value'], r['ms_per_step'], r['cpu_baseline']['cores'])"
