#!/bin/bash
# ODE kernel check: the ODE / golden / pipeline tests, world-2 tests, ODE lines of the batch sweep, per-phase cycles
TAG=${1:-ode}
OUT=gpurun_out; mkdir -p $OUT
(timeout 900 python -m pytest tests/test_gpu_tc_teams.py tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -k "ode or ODE or golden or pipeline or track" 2>&1 | tail -4)
timeout 120 python tools/tc_stress.py 6 f16x2 2>&1 | tail -3
timeout 600 python tools/tc_batch_sweep.py 100 64,256,378 f16x2 0 2>&1 | grep ode | tee $OUT/${TAG}_ode_sweep.txt
timeout 120 python tools/tc_ode_phase_times.py 2>&1 | tee $OUT/${TAG}_tc_ode_phase_cycles.txt
