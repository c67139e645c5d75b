#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== failing / new tests in detail"
timeout 900 python -m pytest tests/test_gpu_tc_teams.py tests/test_gpu_parity.py tests/test_dropin.py tests/test_gpu_tc.py -m gpu -q -x --deselect tests/test_gpu_tc_teams.py::test_tensor_core_samplers_match_reference_goldens 2>&1 | tail -30 | tee $OUT/r2d_pytest_a.log
timeout 600 python -m pytest tests/test_gpu_tc_teams.py -m gpu -q -k "match_reference_goldens" 2>&1 | grep -E "Mismatch|Max abs|Max rel|FAILED|passed|failed|Error" | head -60 | tee $OUT/r2d_pytest_goldens.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden" 2>&1 | tail -15 | tee $OUT/r2d_pytest_parity_goldens.log
echo "== encoder timing (fps levels 2-3 beside level 1)"
timeout 120 python tools/encoder_timing.py 2>&1 | tail -8 | tee $OUT/r2d_encoder_timing.txt
echo "== bench default"
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r2d_bench_c2.json 2> $OUT/r2d_bench_c2.err; tail -c 3000 $OUT/r2d_bench_c2.json
