#!/bin/bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 3 -c 1 \
      -o $OUT/prof_rollout_$TAG -f python bench.py --steps 240 --warmup 40 --no-cpu-baseline > $OUT/ncu_rollout_$TAG.log 2>&1
tail -n 3 $OUT/ncu_rollout_$TAG.log
