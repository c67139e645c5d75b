#!/bin/bash
# ncu evidence for the current kernels: launch list of the bench command + one full capture of the tensor-core PC sampler.
# GPB_PROFILE_NO_COOP=1: Nsight Compute cannot launch a clustered AND cooperative kernel (profiles/README.md).
# usage: tools/gpu_ncu.sh <tag>
TAG=${1:-ncu}
OUT=gpurun_out
mkdir -p $OUT
export GPB_PROFILE_NO_COOP=1
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_launches_bench.log 2>&1
timeout 80 ncu --set full --clock-control none --import-source on -k regex:tc_pc_sampler -s 1 -c 1 -o $OUT/${TAG}_prof_tc_sampler \
    python tools/profile_target.py tc_sampler > $OUT/${TAG}_prof_tc_sampler.log 2>&1
unset GPB_PROFILE_NO_COOP
GPB_SUMMARY_DIR=$OUT python tools/summarize_ncu.py ${TAG} 2>&1 | tail -3
head -8 $OUT/${TAG}_launches_summary.csv
grep -E "gpu__time_duration|tensor|dram__bytes|registers" $OUT/${TAG}_ncu_tc_sampler.csv | head
ls -la $OUT | grep ${TAG}
