#!/bin/bash
# ncu --set full of one steady-state launch of each kernel matching a regex: tools/gpu_ncu_full.sh <tag> <regex> [B] [prec] [skip]
TAG=${1:-x}; RE=${2:-styl_rows3}; B=${3:-64}; PREC=${4:-bf16}; SKIP=${5:-20}
mkdir -p gpurun_out
DIAG_B=$B DIAG_LANES=1 ncu --set full --clock-control none --cache-control none --import-source on -k regex:$RE --launch-skip $SKIP -c 1 \
    -f -o gpurun_out/full_${TAG} python tools/diag_step.py $PREC 1 > gpurun_out/full_${TAG}.log 2>&1
ncu -i gpurun_out/full_${TAG}.ncu-rep --page details --csv > gpurun_out/full_${TAG}_details.csv 2>/dev/null
tail -2 gpurun_out/full_${TAG}.log
