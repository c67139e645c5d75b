#!/bin/bash
# Round 2, second GPU visit: f16x2 arithmetic, tile teams of 4 / 2 / 1, the 256-object batch.
OUT=gpurun_out; mkdir -p $OUT
echo "== 1. f16 two-product MMA"
(timeout 120 python -m pytest tests/test_gpu_tc_teams.py -k two_product_mma -x -q 2>&1 | tail -5) | tee $OUT/r2b_f16_mma.log
echo "== 2. stress f16x2 (team 4)"
timeout 120 python tools/tc_stress.py 20 f16x2 2>&1 | tail -8 | tee $OUT/r2b_stress_f16x2.txt
echo "== 3. team / f16x2 / 256-object tests"
(timeout 900 python -m pytest tests/test_gpu_tc_teams.py -q -s 2>&1 | tail -40) | tee $OUT/r2b_pytest_teams.log
echo "== 4. shipped tc tests"
(timeout 600 python -m pytest tests/test_gpu_tc.py -q 2>&1 | tail -8) | tee $OUT/r2b_pytest_tc.log
echo "== 5. batch sweep"
timeout 600 python tools/tc_batch_sweep.py 100 64,128,256,378 bf16x3,f16x2 0 2>&1 | tee $OUT/r2b_batch_sweep.txt
timeout 300 python tools/tc_batch_sweep.py 100 64 bf16x3,f16x2 1,2,4 2>&1 | tee $OUT/r2b_batch_sweep_teams64.txt
echo "== 6. phase cycles"
timeout 90 python tools/tc_phase_times.py 100 0 f16x2 > $OUT/r2b_phase_f16x2_team4.txt 2>&1; head -32 $OUT/r2b_phase_f16x2_team4.txt
timeout 90 python tools/tc_phase_times.py 100 0 f16x2 1 256 > $OUT/r2b_phase_f16x2_team1_256.txt 2>&1; head -32 $OUT/r2b_phase_f16x2_team1_256.txt
timeout 90 python tools/tc_phase_times.py 100 0 bf16x3 1 256 > $OUT/r2b_phase_bf16x3_team1_256.txt 2>&1; head -32 $OUT/r2b_phase_bf16x3_team1_256.txt
echo "== 7. bench"
for P in bf16x3 f16x2; do
  (timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision $P 2>&1 | tail -1) > $OUT/r2b_bench_c2_$P.json
  python - <<PY
import json
try:
    d = json.load(open("$OUT/r2b_bench_c2_$P.json"))
    print("$P", "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "sampler ms", round(d["roofline"]["kernel_ms"], 3), "e2e", round(d["e2e"]["value"]), "sat", d["roofline"].get("saturating_batch"))
except Exception as e:
    print("$P: no bench line:", e)
PY
done
