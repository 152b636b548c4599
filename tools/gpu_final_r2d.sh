#!/bin/bash
# round 2, final tree: smoke(), then the ncu evidence of the bench command (launch list, DRAM bytes per launch, --set full of the marching kernels)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
bash tools/gpu_launches_r2.sh
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv -k "regex:lines2_|poisson_team" --launch-skip 66 --launch-count 22 --log-file gpurun_out/dram_bench_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_under_ncu_dram_r02.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:lines2_march" --launch-skip 30 --launch-count 10 -f -o /tmp/prof_march_r02 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_march_r02.log 2>&1; tail -2 gpurun_out/ncu_full_march_r02.log
ncu -i /tmp/prof_march_r02.ncu-rep --page raw --csv > gpurun_out/ncu_full_march_r02_raw.csv 2>/dev/null
ls -la gpurun_out/ncu_full_march_r02_raw.csv gpurun_out/dram_bench_r02.csv gpurun_out/launches_bench_r02.csv
