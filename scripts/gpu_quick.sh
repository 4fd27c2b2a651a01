#!/bin/bash
# quick GPU visit: parity tests + bench (no ncu)
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
( timeout 600 python bench.py --no-cpu-baseline ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -n 4 $OUT/pytest_gpu_$TAG.log
cat $OUT/bench_$TAG.json
tail -n 5 $OUT/bench_$TAG.err
