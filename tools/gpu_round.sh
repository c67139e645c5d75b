#!/bin/bash
# One GPU-box visit: tests, smoke, bench (both arms), ncu launch list + full captures.  Everything lands in gpurun_out/.
# usage: tools/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt
nproc >> $OUT/${TAG}_gpu.txt
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $OUT/${TAG}_pytest_gpu.log; tail -5 $OUT/${TAG}_pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.log
echo "== bench config 2" ; timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee $OUT/${TAG}_bench_c2.json
echo "== bench config 3" ; timeout 600 python bench.py --steps 5 --warmup 3 --config 3 --no-cpu-baseline 2>&1 | tail -3 | tee $OUT/${TAG}_bench_c3.json
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee $OUT/${TAG}_bench_ref.json
echo "== bench ode (extra line)" ; timeout 600 python bench.py --sampler ode --steps 10 --warmup 3 2>&1 | tail -1 | tee $OUT/${TAG}_bench_ode.json
# Nsight Compute cannot launch a kernel that is both clustered and cooperative: GPB_PROFILE_NO_COOP=1 drops the cooperative
# attribute for the profiling runs only (profiles/README.md); numbers printed under ncu are never bench values.
export GPB_PROFILE_NO_COOP=1
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_launches_bench.log 2>&1
echo "== ncu full: sampler"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_pc_sampler -s 1 -c 1 -o $OUT/${TAG}_prof_tc_sampler \
    python tools/profile_target.py tc_sampler > $OUT/${TAG}_prof_tc_sampler.log 2>&1
echo "== ncu full: ode sampler"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_ode_sampler -s 1 -c 1 -o $OUT/${TAG}_prof_tc_ode_sampler \
    python tools/profile_target.py tc_ode > $OUT/${TAG}_prof_tc_ode_sampler.log 2>&1
echo "== ncu full: encoder"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sa_kernel|sa_small_tc|sa3_tc|ga_gemm|fps3|point_gemm|object_bias' -s 15 -c 15 -o $OUT/${TAG}_prof_encoder \
    python tools/profile_target.py encoder > $OUT/${TAG}_prof_encoder.log 2>&1
unset GPB_PROFILE_NO_COOP
GPB_SUMMARY_DIR=$OUT python tools/summarize_ncu.py ${TAG} 2>&1 | tail -5
ls -la $OUT | tail -20
