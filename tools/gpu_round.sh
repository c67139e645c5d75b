#!/bin/bash
# One GPU-box visit: tests, smoke (plain and under ncu), bench (both arms), ncu launch list + full captures.  Everything lands in
# gpurun_out/; tools/summarize_ncu.py <tag> then writes the tracked summaries under profiles/ (run it in the build container).
# usage: tools/gpu_round.sh <tag> [skip-tests]
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt
nproc >> $OUT/${TAG}_gpu.txt
if [ "$2" != "skip-tests" ]; then
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $OUT/${TAG}_pytest_gpu.log; tail -5 $OUT/${TAG}_pytest_gpu.log
fi
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.log
echo "== smoke under ncu (launch list: the tensor-core samplers must appear)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_smoke_launches.csv \
    python __graft_entry__.py smoke > $OUT/${TAG}_smoke_ncu.log 2>&1; echo "ncu rc=$?"; grep -c "tc_pc_sampler\|tc_ode_sampler" $OUT/${TAG}_smoke_launches.csv
echo "== bench (default line: config 2 + extra keys)" ; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee $OUT/${TAG}_bench_c2.json
echo "== bench config 3" ; timeout 600 python bench.py --steps 5 --warmup 3 --config 3 --no-cpu-baseline --no-extras 2>&1 | tail -3 | tee $OUT/${TAG}_bench_c3.json
echo "== bench bf16x3" ; timeout 600 python bench.py --steps 5 --warmup 3 --precision bf16x3 --no-cpu-baseline --no-extras 2>&1 | tail -1 | tee $OUT/${TAG}_bench_c2_bf16x3.json
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee $OUT/${TAG}_bench_ref.json
echo "== bench ode (own line)" ; timeout 600 python bench.py --sampler ode --steps 10 --warmup 3 2>&1 | tail -1 | tee $OUT/${TAG}_bench_ode.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $OUT/${TAG}_launches_bench.log 2>&1
echo "== ncu full: sampler (bench shape, auto precision)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_pc_sampler -s 1 -c 1 -o $OUT/${TAG}_prof_tc_sampler \
    python tools/profile_target.py tc_sampler 500 f16x2 64 $OUT/${TAG}_prof_tc_sampler.shape.json > $OUT/${TAG}_prof_tc_sampler.log 2>&1
echo "== ncu full: sampler, one tile per SM (378 objects)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_pc_sampler -s 1 -c 1 -o $OUT/${TAG}_prof_tc_sampler_sat \
    python tools/profile_target.py tc_sampler 100 f16x2 378 $OUT/${TAG}_prof_tc_sampler_sat.shape.json > $OUT/${TAG}_prof_tc_sampler_sat.log 2>&1
echo "== ncu full: ode sampler"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_ode_sampler -s 1 -c 1 -o $OUT/${TAG}_prof_tc_ode_sampler \
    python tools/profile_target.py tc_ode 500 f16x2 64 $OUT/${TAG}_prof_tc_ode_sampler.shape.json > $OUT/${TAG}_prof_tc_ode_sampler.log 2>&1
echo "== ncu full: encoder"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sa_kernel|sa_small_tc|sa3_tc|ga_gemm|fps3|point_gemm|object_bias' -s 15 -c 15 -o $OUT/${TAG}_prof_encoder \
    python tools/profile_target.py encoder > $OUT/${TAG}_prof_encoder.log 2>&1
echo "== ncu full: energy + rank/pool"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'energy|rank_pool|trunk_eval' -s 2 -c 4 -o $OUT/${TAG}_prof_energy_rank_pool \
    python tools/profile_target.py energy_rank_pool > $OUT/${TAG}_prof_energy_rank_pool.log 2>&1
# summaries are made HERE (the box has ncu): only the small CSVs travel back (gpurun_out is capped at 64 MiB), plus the per-instruction
# source page of the dominant kernel; the .ncu-rep files stay on the box
GPB_SUMMARY_DIR=$OUT python tools/summarize_ncu.py ${TAG} 2>&1 | tail -12
for n in tc_sampler tc_sampler_sat; do
  ncu -i $OUT/${TAG}_prof_$n.ncu-rep --page source --csv > $OUT/${TAG}_source_$n.csv 2>/dev/null
  gzip -f $OUT/${TAG}_source_$n.csv
done
rm -f $OUT/*.ncu-rep
ls -la $OUT | tail -40; du -sh $OUT
