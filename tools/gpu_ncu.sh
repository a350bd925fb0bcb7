#!/bin/bash
# usage (under gpurun): bash tools/gpu_ncu.sh <tag> <kernel regex> [workload] [extra env assignments...]
TAG=$1; KRN=$2; WL=${3:-c3}; shift 3
mkdir -p gpurun_out
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRN -s ${NCU_SKIP:-2} -c 1 -f -o gpurun_out/${TAG}_ncu \
  python bench.py --workload $WL --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
ls -la gpurun_out/${TAG}_ncu.ncu-rep
