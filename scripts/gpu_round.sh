#!/bin/bash
# One gpurun visit: smoke, GPU parity tests, bench (own + reference arm), ncu launch list and full captures.
# usage (from the repo root, on the GPU box): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi > $OUT/nvsmi_$TAG.txt 2>&1
lscpu | head -20 > $OUT/lscpu_$TAG.txt 2>&1
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke_$TAG.log 2>&1
echo "smoke exit $?" >> $OUT/smoke_$TAG.log
( timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
( timeout 600 python bench.py ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
( timeout 600 python bench.py --impl reference --steps 200 --warmup 10 ) > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
# launch list of the bench command (cold-cache, serialised: shares only)
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 160 --warmup 40 --no-cpu-baseline ) > $OUT/ncu_launches_$TAG.log 2>&1
# full capture of the two kernels
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 100 -c 3 \
    -o $OUT/prof_step_$TAG -f python bench.py --steps 80 --warmup 40 --no-cpu-baseline ) > $OUT/ncu_step_$TAG.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gen -s 100 -c 3 \
    -o $OUT/prof_gen_$TAG -f python bench.py --steps 80 --warmup 40 --no-cpu-baseline ) > $OUT/ncu_gen_$TAG.log 2>&1
tail -n 3 $OUT/smoke_$TAG.log $OUT/pytest_gpu_$TAG.log
cat $OUT/bench_$TAG.json
tail -n 5 $OUT/bench_$TAG.err
