#!/bin/bash
# Per-kernel durations (warm caches) of rg_denoise at one batch size: tools/gpu_kernels.sh <tag> [B] [precision]
TAG=${1:-x}; B=${2:-64}; PREC=${3:-bf16}
mkdir -p gpurun_out
DIAG_B=$B DIAG_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv \
    --log-file gpurun_out/kern_${TAG}.csv python tools/diag_step.py $PREC 1 > gpurun_out/kern_${TAG}.log 2>&1
tail -2 gpurun_out/kern_${TAG}.log
