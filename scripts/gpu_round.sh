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
# full capture of the dominant kernel (fused k_rollout: two stepping warps + the generator warp per CTA)
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 3 -c 1 \
    -o $OUT/prof_rollout_$TAG -f python bench.py --steps 240 --warmup 40 --no-cpu-baseline ) > $OUT/ncu_rollout_$TAG.log 2>&1
# A/B: refill passes instead of the fused generator warp; fast-consuming levels both ways
for lv in GoToLocal GoToObjS4 GoToLocalS5N2 GoToObjS6; do for f in 1 0; do
  echo -n "$lv fused=$f: " >> $OUT/ab_$TAG.log
  BB_GEN_FUSED=$f timeout 300 python bench.py --no-cpu-baseline --steps 2000 --warmup 200 --level $lv 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.3e ms_per_step %.4f rollout %.1f us refill/launch %.1f us errors %d' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']*1e3, d['roofline']['refill_ms_per_launch']*1e3, d['counters']['errors']))
" >> $OUT/ab_$TAG.log
done; done
cat $OUT/ab_$TAG.log
tail -n 3 $OUT/smoke_$TAG.log $OUT/pytest_gpu_$TAG.log
cat $OUT/bench_$TAG.json
tail -n 5 $OUT/bench_$TAG.err
