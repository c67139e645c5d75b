#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_ext.py -m gpu -q -x 2>&1 | tail -3)
echo "== serial schedule"; GPB_ENC_SERIAL=1 timeout 120 python tools/encoder_timing.py 2>&1 | grep "bf16x3\|max" | tee $OUT/r2i_encoder_serial.txt
echo "== forked schedule"; timeout 120 python tools/encoder_timing.py 2>&1 | grep "bf16x3\|max" | tee $OUT/r2i_encoder_forked.txt
for B in 8 256; do echo "B=$B"; GPB_ENC_SERIAL=1 timeout 120 python tools/encoder_timing.py $B 2>&1 | grep "bf16x3"; timeout 120 python tools/encoder_timing.py $B 2>&1 | grep "bf16x3"; done | tee $OUT/r2i_encoder_sizes.txt
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'sampler',round(d['roofline']['kernel_ms'],3),'e2e',round(d['e2e']['value']),'pipelined',round(d['pipelined']['value']))"
