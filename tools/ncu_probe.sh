#!/bin/bash
# Nsight Compute on the cluster + cooperative sampler kernel: only possible with the cooperative attribute off.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1b}
export GPB_PROFILE_NO_COOP=1
echo "== launch list of one bench step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_launches_bench.log 2>&1; echo rc=$?; tail -2 $OUT/${TAG}_launches.csv | cut -c1-300
echo "== full set, tc sampler"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_pc_sampler -s 1 -c 1 -o $OUT/${TAG}_prof_tc_sampler \
    python tools/profile_target.py tc_sampler > $OUT/${TAG}_prof_tc_sampler.log 2>&1; echo rc=$?; tail -3 $OUT/${TAG}_prof_tc_sampler.log
