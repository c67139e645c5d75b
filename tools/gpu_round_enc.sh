#!/bin/bash
# GPU visit focused on the encoder.  usage: tools/gpu_round_enc.sh <tag>
TAG=${1:-enc1}
OUT=gpurun_out
mkdir -p $OUT
echo "== encoder tests"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "encoder or golden" 2>&1 | tail -30 > $OUT/${TAG}_pytest_enc.log; tail -8 $OUT/${TAG}_pytest_enc.log
echo "== encoder timing"
timeout 200 python tools/encoder_timing.py 2>&1 | tail -6 | tee $OUT/${TAG}_enc_timing.txt
echo "== launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_enc_launches.csv python tools/profile_target.py encoder > /dev/null 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("$OUT/${TAG}_enc_launches.csv") if l.startswith('"')))
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: agg.setdefault(r[ki][:60], []).append(float(r[vi].replace(",", "")))
    except Exception: pass
for k, v in agg.items(): print(f"{k:62s} n={len(v)} avg_us={sum(v)/len(v)/1000:.1f}")
PY
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/${TAG}_bench.json
