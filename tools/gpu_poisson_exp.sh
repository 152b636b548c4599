#!/bin/bash
# Poisson y-solve variants at C3 (team kernel: kx per CTA, prefetch distance), thread-per-mode kernel for comparison, C4-like ny = 1024
for cfg in "poisson_warp=4" "poisson_warp=4,poisson_pf=0" "poisson_warp=4,poisson_pf=592" "poisson_warp=8" "poisson_warp=0"; do
  echo "== $cfg"
  POISSON_IL=0 POISSON_CFG=3:0 python tools/bench_ops.py --poisson --no-lines --shape 1024,512,1024 --iters 5 --warmup 2 --tune $cfg 2>&1 | tail -2
done
for cfg in "poisson_warp=4" "poisson_warp=2"; do
  echo "== ny=1024 (256,1024,1024) $cfg"
  POISSON_IL=0 POISSON_CFG=3:0 python tools/bench_ops.py --poisson --no-lines --shape 256,1024,1024 --iters 5 --warmup 2 --tune $cfg 2>&1 | tail -2
done
