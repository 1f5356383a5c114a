#!/bin/bash
# bench (+ optional ncu launch list / full capture of a kernel regex).  bash tools/gpu_bench.sh TAG PRECISION [KERNEL_REGEX]
TAG=${1:-r01}; PREC=${2:-bf16}; KREG=${3:-}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --precision $PREC 2>gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json
tail -5 gpurun_out/bench_$TAG.err
if [ -n "$KREG" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1500 --csv \
      --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 1 --knn-n 20000 --precision $PREC \
      > gpurun_out/ncu_list_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:$KREG -s 600 -c 4 -f \
      -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 1 --cpu-seconds 1 --knn-n 20000 --precision $PREC \
      > gpurun_out/ncu_full_$TAG.log 2>&1
  ls -la gpurun_out | tail -8
fi
