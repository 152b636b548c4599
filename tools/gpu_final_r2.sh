#!/bin/bash
# round 2: the GPU suite, the bench line, and the ncu evidence of the same command (launch list, DRAM bytes per launch,
# --set full of the dominant kernels)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_gpu_suite.log
timeout 500 python bench.py > gpurun_out/bench_1gpu_c3_r02.json 2> gpurun_out/bench_1gpu_c3_r02.err; tail -c 300 gpurun_out/bench_1gpu_c3_r02.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; tail -c 300 gpurun_out/bench_ref_r02.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_under_ncu_r02.log 2>&1
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv -k "regex:lines2_|poisson_team" --launch-skip 66 --launch-count 22 --log-file gpurun_out/dram_bench_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_under_ncu_dram_r02.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:lines2_march|lines2_contig|lines2_strided|poisson_team" --launch-skip 66 --launch-count 22 -f -o gpurun_out/prof_bench_r02 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_r02.log 2>&1; tail -2 gpurun_out/ncu_full_r02.log
python -c "
import json;d=json.load(open('gpurun_out/bench_1gpu_c3_r02.json'));print(round(d['value'],3),round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items()}); print(d['cpu_baseline']); print(d['e2e'])
r=json.load(open('gpurun_out/bench_ref_r02.json')); print('reference arm', r['value'], r['ms_per_step'], r['cpu_baseline']['cores'])"
