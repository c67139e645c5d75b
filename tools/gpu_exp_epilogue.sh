#!/bin/bash
# GPU-box visit for the compile-time epilogue experiments: build the variants BEFORE the visit with
#   make -C genpose_b200/csrc variant NAME=rz   EXTRA=-DGPB_EPI_RZ_RELU=1
#   make -C genpose_b200/csrc variant NAME=ld2  EXTRA=-DGPB_EPI_LD2=1
#   make -C genpose_b200/csrc variant NAME=both EXTRA="-DGPB_EPI_RZ_RELU=1 -DGPB_EPI_LD2=1"
# then, per variant: small-grid stress (bit-identical, <= 1e-3 from the FFMA kernel), the tensor-core GPU tests, the bench line.
# usage: tools/gpu_exp_epilogue.sh <tag> [variant names...]
TAG=${1:-epi}; shift
OUT=gpurun_out
mkdir -p $OUT
for V in default "$@"; do
  if [ "$V" = default ]; then unset GPB_LIB; else export GPB_LIB=$PWD/genpose_b200/libgenpose_b200_$V.so; fi
  echo "== variant $V (${GPB_LIB:-default library})"
  timeout 60 python tools/tc_stress.py 30 2>&1 | tail -2
  (timeout 200 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2)
  (timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > $OUT/${TAG}_bench_c2_$V.json
  python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_c2_$V.json"))
    print("$V", "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "sampler ms", round(d["roofline"]["kernel_ms"], 3))
except Exception as e:
    print("$V: no bench line:", e)
PY
  timeout 60 python tools/tc_phase_times.py 100 > $OUT/${TAG}_tc_phase_cycles_$V.txt 2>&1
done
unset GPB_LIB
