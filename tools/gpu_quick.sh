#!/bin/bash
# quick kernel-change check: stress, tensor-core tests, phase cycles, batch sweep
TAG=${1:-q}
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python tools/tc_stress.py 20 f16x2 2>&1 | tail -3
timeout 120 python tools/tc_stress.py 10 bf16x3 2>&1 | tail -2
(timeout 900 python -m pytest tests/test_gpu_tc_teams.py tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -4) | tee $OUT/${TAG}_pytest_tc.log
timeout 90 python tools/tc_phase_times.py 100 0 f16x2 > $OUT/${TAG}_phase_f16x2_team4.txt 2>&1; head -29 $OUT/${TAG}_phase_f16x2_team4.txt | tail -14
timeout 90 python tools/tc_phase_times.py 100 0 f16x2 1 256 > $OUT/${TAG}_phase_f16x2_team1_256.txt 2>&1; head -29 $OUT/${TAG}_phase_f16x2_team1_256.txt | tail -14
timeout 600 python tools/tc_batch_sweep.py 100 64,256,378 f16x2 0 2>&1 | tee $OUT/${TAG}_batch_sweep.txt
