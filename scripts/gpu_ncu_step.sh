#!/bin/bash
# ncu --set full of the step kernel variants: $1 = tag
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
for k in cols lane; do
  BB_STEP_KERNEL=$k timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 100 -c 4 \
      -o $OUT/prof_${k}_$TAG -f python bench.py --steps 80 --warmup 40 --no-cpu-baseline > $OUT/ncu_${k}_$TAG.log 2>&1
done
BB_STEP_KERNEL=lane timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 100 -c 4 \
      -o $OUT/prof_lane8k_$TAG -f python bench.py --steps 80 --warmup 40 --no-cpu-baseline --envs 8192 > $OUT/ncu_lane8k_$TAG.log 2>&1
ls -la $OUT/*.ncu-rep | tail -5
