#!/bin/bash
# compute-sanitizer memcheck over short PC + ODE launches of every tile-team size and both arithmetics (twelve sampler instantiations)
TAG=${1:-mc}
OUT=gpurun_out; mkdir -p $OUT
for P in f16x2 bf16x3; do
  echo "== memcheck $P (teams 1, 2, 4; PC T=6 + ODE T0=0.15)"
  timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/mc_small.py 1,2,4 $P 2>&1 | grep -v "^=========     \|Host Frame\|Device Frame" | tail -12
done | tee $OUT/${TAG}_memcheck.txt
