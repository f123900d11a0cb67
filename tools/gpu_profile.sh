#!/bin/bash
# Runs on the GPU box (via gpurun): plain bench first, then the ncu launch list and one
# full capture of the dominant kernel for the same command line (B200_PROFILING.md recipe).
#   tools/gpu_profile.sh <tag> <kernel-regex> [bench args...]
set -u
TAG=$1; KRE=$2; shift 2
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline $*"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -f -o $OUT/${TAG}_prof $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
tail -2 $OUT/${TAG}_plain.log
