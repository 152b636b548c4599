#!/bin/bash
# round 2, final state: remaining GPU test modules, the bench line, the reference arm, and the ncu evidence of the same command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_case10_gpu.py tests/test_fourier_gpu.py tests/test_fullsize_gpu.py tests/test_poisson_gpu.py tests/test_solvers_gpu.py tests/test_abi.py tests/test_cpu_baseline.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2_gpu_suite_rest.log
timeout 500 python bench.py > gpurun_out/bench_1gpu_c3_r02.json 2> gpurun_out/bench_1gpu_c3_r02.err; tail -c 300 gpurun_out/bench_1gpu_c3_r02.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; tail -c 300 gpurun_out/bench_ref_r02.err
bash tools/gpu_launches_r2.sh
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv -k "regex:lines2_|poisson_team" --launch-skip 66 --launch-count 22 --log-file gpurun_out/dram_bench_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_under_ncu_dram_r02.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:lines2_march|lines2_contig|lines2_strided|poisson_team" --launch-skip 66 --launch-count 22 -f -o /tmp/prof_bench_r02 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_r02.log 2>&1; tail -2 gpurun_out/ncu_full_r02.log
python tools/ncu_key.py /tmp/prof_bench_r02.ncu-rep > gpurun_out/ncu_full_bench_r02.txt 2>&1
ncu -i /tmp/prof_bench_r02.ncu-rep --page raw --csv > gpurun_out/ncu_full_bench_r02_raw.csv 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/bench_1gpu_c3_r02.json'));print(round(d['value'],3),round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['breakdown_ms'].items()}); print(d['cpu_baseline']); print(d['e2e']); print(d['roofline'])
r=json.load(open('gpurun_out/bench_ref_r02.json')); print('reference arm', r['value'], r['ms_per_step'], r['cpu_baseline']['cores'])"
