#!/bin/bash
# Last GPU-box visit of a round: stress the small-grid sampler first (a hang there ends the visit early), then the GPU suite,
# smoke, and the two bench lines.  usage: tools/gpu_final.sh <tag>
TAG=${1:-fin}
OUT=gpurun_out
mkdir -p $OUT
timeout 60 python tools/tc_stress.py 60 > $OUT/${TAG}_tc_stress.txt 2>&1; RC=$?; echo "stress rc=$RC" >> $OUT/${TAG}_tc_stress.txt
cat $OUT/${TAG}_tc_stress.txt
if [ $RC -ne 0 ]; then echo "stress failed: stopping"; exit 1; fi
(timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $OUT/${TAG}_pytest_gpu.log; tail -3 $OUT/${TAG}_pytest_gpu.log
(timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > $OUT/${TAG}_bench_c2.json
(timeout 100 python __graft_entry__.py smoke 2>&1 | tail -4) > $OUT/${TAG}_smoke.log; cat $OUT/${TAG}_smoke.log
(timeout 100 python bench.py --steps 5 --warmup 3 --config 3 --no-cpu-baseline 2>&1 | tail -1) > $OUT/${TAG}_bench_c3.json
timeout 60 python tools/tc_phase_times.py 100 > $OUT/${TAG}_tc_phase_cycles.txt 2>&1
python - <<PY
import json
for c in ("c2", "c3"):
    try:
        d = json.load(open("$OUT/${TAG}_bench_%s.json" % c))
        print(c, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "sampler ms", round(d["roofline"]["kernel_ms"], 3), "e2e", round(d["e2e"]["value"]),
              "pipelined", round(d.get("pipelined", {}).get("value", 0)), "clocks", d["clocks"])
    except Exception as e:
        print(c, "no bench line:", e)
PY
