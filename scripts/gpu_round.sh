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
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 600 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 160 --warmup 40 --no-cpu-baseline ) > $OUT/ncu_launches_$TAG.log 2>&1
# full capture of the two kernels
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 3 -c 1 \
    -o $OUT/prof_rollout_$TAG -f python bench.py --steps 240 --warmup 40 --no-cpu-baseline ) > $OUT/ncu_rollout_$TAG.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gen_small -s 4 -c 1 \
    -o $OUT/prof_gensmall_$TAG -f python bench.py --steps 240 --warmup 40 --no-cpu-baseline ) > $OUT/ncu_gensmall_$TAG.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step8 -s 50 -c 2 \
    -o $OUT/prof_step8_$TAG -f python bench.py --steps 240 --warmup 40 --no-cpu-baseline ) > $OUT/ncu_step8_$TAG.log 2>&1
tail -n 3 $OUT/smoke_$TAG.log $OUT/pytest_gpu_$TAG.log
cat $OUT/bench_$TAG.json
tail -n 5 $OUT/bench_$TAG.err
