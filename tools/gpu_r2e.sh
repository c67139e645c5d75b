#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $OUT/r2e_pytest_gpu.log; tail -12 $OUT/r2e_pytest_gpu.log
echo "== undamped golden numbers"
timeout 300 python -m pytest tests/test_gpu_tc_teams.py -m gpu -q -s -k undamped 2>&1 | grep -E "kappa|passed|failed" | tee $OUT/r2e_undamped.txt
echo "== phase cycles (packed epilogues)"
timeout 90 python tools/tc_phase_times.py 100 0 f16x2 > $OUT/r2e_phase_f16x2_team4.txt 2>&1; head -30 $OUT/r2e_phase_f16x2_team4.txt
timeout 90 python tools/tc_phase_times.py 100 0 f16x2 1 256 > $OUT/r2e_phase_f16x2_team1_256.txt 2>&1; head -30 $OUT/r2e_phase_f16x2_team1_256.txt
echo "== batch sweep"
timeout 600 python tools/tc_batch_sweep.py 100 64,256,378 f16x2 0 2>&1 | tee $OUT/r2e_batch_sweep.txt
echo "== encoder timing"
timeout 120 python tools/encoder_timing.py 2>&1 | tail -8 | tee $OUT/r2e_encoder_timing.txt
echo "== bench default"
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r2e_bench_c2.json 2> $OUT/r2e_bench_c2.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2e_bench_c2.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "sampler", round(d["roofline"]["kernel_ms"], 3), "frac", round(d["roofline"]["frac"], 4), "e2e", round(d["e2e"]["value"]))
print("pipelined", d.get("pipelined", {}).get("value"))
print("config3", {k: d["config3"][k] for k in ("value", "ms_per_step", "kernel_ms")}, d["config3"]["e2e"]["value"])
print("ode", {k: d["ode_recipe"][k] for k in ("value", "ms_per_step", "nfev")}, d["ode_recipe"]["roofline"]["kernel_ms"], d["ode_recipe"]["roofline"]["frac"])
print("sat", d["roofline"].get("saturating_batch"))
print("gpu_torch", d.get("gpu_torch_baseline"))
print("cpu", d.get("cpu_baseline"))
print("clocks", d.get("clocks"))
PY
