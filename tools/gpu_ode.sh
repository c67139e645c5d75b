#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
(timeout 900 python -m pytest tests/test_gpu_tc_teams.py tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -k "ode or ODE or golden or pipeline or track" 2>&1 | tail -4)
(timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q 2>&1 | tail -4)
timeout 600 python tools/tc_batch_sweep.py 100 64,256 f16x2,bf16x3 0 2>&1 | grep ode | tee $OUT/r2j_ode_sweep.txt
timeout 120 python tools/tc_ode_phase_times.py 2>&1 | tee $OUT/r2j_tc_ode_phase_cycles.txt
