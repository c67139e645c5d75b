#!/bin/bash
# First GPU-box visit for the experimental two-product tensor-core samplers (precision='bf16x2').
# Order: the cheapest, most decisive checks first; every step under its own timeout (a hang is a failure, not a strike).
# usage: tools/gpu_exp_bf16x2.sh <tag>
TAG=${1:-x2}
OUT=gpurun_out
mkdir -p $OUT
export GPB_EXPERIMENTAL=1
echo "== 1. mixed-format instruction (A = bf16, B = fp16)"
(timeout 60 python -m pytest tests/test_gpu_tc16.py -k mixed_format -x -q 2>&1 | tail -15) | tee $OUT/${TAG}_mixed_format.log
echo "== 2. small-grid stress, two-product PC sampler"
timeout 60 python tools/tc_stress.py 30 bf16x2 2>&1 | tail -8 | tee $OUT/${TAG}_stress.txt
echo "== 3. samplers against oracle / FFMA kernel / three-product kernel"
(timeout 200 python -m pytest tests/test_gpu_tc16.py -x -q -s 2>&1 | tail -30) | tee $OUT/${TAG}_pytest_tc16.log
echo "== 4. bench, both precisions"
for P in bf16x3 bf16x2; do
  (timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision $P 2>&1 | tail -1) > $OUT/${TAG}_bench_c2_$P.json
  python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_c2_$P.json"))
    print("$P", "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "sampler ms", round(d["roofline"]["kernel_ms"], 3), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$P: no bench line:", e)
PY
done
echo "== 5. per-phase cycles"
timeout 60 python tools/tc_phase_times.py 100 0 bf16x2 > $OUT/${TAG}_tc_phase_cycles_bf16x2.txt 2>&1; head -28 $OUT/${TAG}_tc_phase_cycles_bf16x2.txt
echo "== 6. shipped path untouched: full GPU suite"
unset GPB_EXPERIMENTAL
(timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) | tee $OUT/${TAG}_pytest_gpu.log
