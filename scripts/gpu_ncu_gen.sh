#!/bin/bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gen_small -s 6 -c 2 \
      -o $OUT/prof_gensmall_$TAG -f python bench.py --steps 400 --warmup 40 --no-cpu-baseline > $OUT/ncu_gensmall_$TAG.log 2>&1
tail -n 2 $OUT/ncu_gensmall_$TAG.log
