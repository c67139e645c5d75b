#!/bin/bash
# last check of a round after a late kernel change: full GPU suite, smoke, memcheck, the two bench lines that name an arithmetic
TAG=${1:-fin}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee $OUT/${TAG}_pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
bash tools/gpu_memcheck.sh ${TAG} | grep -v "^========= COMPUTE" 
echo "== bench c2"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > $OUT/${TAG}_bench_c2.json; cut -c1-330 $OUT/${TAG}_bench_c2.json
echo "== bench bf16x3"; timeout 600 python bench.py --steps 5 --warmup 3 --precision bf16x3 --no-cpu-baseline --no-extras 2>&1 | tail -1 > $OUT/${TAG}_bench_c2_bf16x3.json; cut -c1-330 $OUT/${TAG}_bench_c2_bf16x3.json
