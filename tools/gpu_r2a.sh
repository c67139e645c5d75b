#!/bin/bash
# Round 2, first GPU visit: the two-product samplers (bf16x2) and the epilogue variants that round 1 left unmeasured.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/r2a_gpu.txt
export GPB_EXPERIMENTAL=1
echo "== 1. mixed-format instruction"
(timeout 120 python -m pytest tests/test_gpu_tc16.py -k mixed_format -x -q 2>&1 | tail -15) | tee $OUT/r2a_mixed_format.log
echo "== 2. stress bf16x2"
timeout 90 python tools/tc_stress.py 30 bf16x2 2>&1 | tail -8 | tee $OUT/r2a_stress_x2.txt
echo "== 3. tc16 tests"
(timeout 300 python -m pytest tests/test_gpu_tc16.py -x -q -s 2>&1 | tail -30) | tee $OUT/r2a_pytest_tc16.log
echo "== 4. bench both"
for P in bf16x3 bf16x2; do
  (timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision $P 2>&1 | tail -1) > $OUT/r2a_bench_c2_$P.json
  python - <<PY
import json
try:
    d = json.load(open("$OUT/r2a_bench_c2_$P.json"))
    print("$P", "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "sampler ms", round(d["roofline"]["kernel_ms"], 3), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$P: no bench line:", e)
PY
done
echo "== 5. phase cycles x2"
timeout 90 python tools/tc_phase_times.py 100 0 bf16x2 > $OUT/r2a_tc_phase_cycles_bf16x2.txt 2>&1; head -40 $OUT/r2a_tc_phase_cycles_bf16x2.txt
unset GPB_EXPERIMENTAL
echo "== 6. epilogue variants"
bash tools/gpu_exp_epilogue.sh r2a rz ld2 both 2>&1 | tail -30
