#!/bin/bash
# GPU visit focused on the tcgen05 path.  usage: tools/gpu_round_tc.sh <tag>
TAG=${1:-tc1}
OUT=gpurun_out
mkdir -p $OUT
echo "== selftest + tc tests"
timeout 300 python -m pytest tests/test_gpu_tc.py -q -s 2>&1 | tail -60 > $OUT/${TAG}_pytest_tc.log; tail -15 $OUT/${TAG}_pytest_tc.log
if grep -q "passed" $OUT/${TAG}_pytest_tc.log && ! grep -q "failed" $OUT/${TAG}_pytest_tc.log; then
  echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $OUT/${TAG}_pytest_gpu.log; tail -4 $OUT/${TAG}_pytest_gpu.log
  echo "== bench auto (tc)"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee $OUT/${TAG}_bench_tc.json
  echo "== bench fp32"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision fp32 2>&1 | tail -2 | tee $OUT/${TAG}_bench_fp32.json
  echo "== bench config3 (tc)"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --config 3 2>&1 | tail -2 | tee $OUT/${TAG}_bench_c3.json
  echo "== ncu tc sampler"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_pc_sampler -s 1 -c 1 -o $OUT/${TAG}_prof_tc_sampler \
      python tools/profile_target.py tc_sampler > $OUT/${TAG}_prof_tc_sampler.log 2>&1
  tail -3 $OUT/${TAG}_prof_tc_sampler.log
else
  echo "== tc tests failed; compute-sanitizer on the selftest"
  timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_tc.py -q -x -k "conventions" 2>&1 | tail -40 > $OUT/${TAG}_sanitizer.log; tail -20 $OUT/${TAG}_sanitizer.log
fi
ls -la $OUT | tail -12
