#!/bin/bash
# Short GPU-box visit for one kernel experiment: GPU tests, per-phase cycle stamps of the PC sampler, config-2 bench line.
# usage: tools/gpu_quick.sh <tag> [pytest args]
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
(timeout 240 python -m pytest tests -m gpu -x -q ${@:2} 2>&1 | tail -15) > $OUT/${TAG}_pytest_gpu.log
tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 120 python tools/tc_phase_times.py 100 > $OUT/${TAG}_tc_phase_cycles.txt 2>&1
cat $OUT/${TAG}_tc_phase_cycles.txt
(timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > $OUT/${TAG}_bench_c2.json
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_c2.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "sampler ms", round(d["roofline"]["kernel_ms"], 3), "e2e", round(d["e2e"]["value"]),
      "pipelined", round(d.get("pipelined", {}).get("value", 0)), "clocks", d["clocks"])
PY
