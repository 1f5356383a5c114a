#!/bin/bash
# One gpurun call: GPU parity tests, smoke, 1-GPU bench, ncu launch list + full capture of the top kernel.
# Usage (from the repo root on the GPU box): bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_$TAG.log
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke_$TAG.log
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json
tail -5 gpurun_out/bench_$TAG.err
if [ "$2" != "noncu" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv \
      --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 1 --knn-n 20000 \
      > gpurun_out/ncu_list_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 400 -c 3 -f \
      -o gpurun_out/prof_gemm_$TAG python bench.py --steps 1 --warmup 1 --cpu-seconds 1 --knn-n 20000 \
      > gpurun_out/ncu_full_$TAG.log 2>&1
  ls -la gpurun_out | tail -12
fi
