#!/bin/bash
# source-level profile of the x Burgers kernel (one launch)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:lines2_contig" --launch-skip 2 --launch-count 1 -f -o /tmp/prof_x python tools/bench_ops.py --shape 1024,512,512 --only "Burgers_X U_IN" --persist 0 --iters 3 --warmup 2 > gpurun_out/ncu_x.log 2>&1
tail -3 gpurun_out/ncu_x.log
ls -la /tmp/prof_x.ncu-rep
ncu -i /tmp/prof_x.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_x_sass.csv 2>/dev/null
ncu -i /tmp/prof_x.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/ncu_x_src.csv 2>/dev/null
ncu -i /tmp/prof_x.ncu-rep --page details > gpurun_out/ncu_x_details.txt 2>/dev/null
ls -la gpurun_out/ncu_x*
