#!/bin/bash
# Round-2 fourth GPU visit: pipelined k_rollout_cta (stepper warp + observer warps), branch-light verifier, concurrent
# generation passes at 2 blocks per SM; compute-sanitizer runs.
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
for lv in GoTo BossLevel; do
  echo "== $lv" >> $OUT/multiroom_$TAG.log
  ( timeout 200 python bench.py --brief --level $lv --envs 32768 --steps 2000 --warmup 200 ) >> $OUT/multiroom_$TAG.log 2>&1
done
for b in 1 4; do
  echo "== BossLevel BB_GEN_BESIDE_BLOCKS_PER_SM=$b" >> $OUT/multiroom_$TAG.log
  ( BB_GEN_BESIDE_BLOCKS_PER_SM=$b timeout 200 python bench.py --brief --level BossLevel --envs 32768 --steps 2000 --warmup 200 ) >> $OUT/multiroom_$TAG.log 2>&1
done
echo "== BossLevel BB_GEN_CONCURRENT=0" >> $OUT/multiroom_$TAG.log
( BB_GEN_CONCURRENT=0 timeout 200 python bench.py --brief --level BossLevel --envs 32768 --steps 2000 --warmup 200 ) >> $OUT/multiroom_$TAG.log 2>&1
for pp in 0 1; do for lv in GoToLocal PickupLoc GoToObjS4; do
  echo "== $lv BB_ROLLOUT_PIPE=$pp" >> $OUT/multiroom_$TAG.log
  ( BB_ROLLOUT_PIPE=$pp timeout 200 python bench.py --brief --level $lv --steps 2000 --warmup 200 ) >> $OUT/multiroom_$TAG.log 2>&1
done; done
echo "== GoToLocal BB_ROLLOUT_KERNEL=cta" >> $OUT/multiroom_$TAG.log
( BB_ROLLOUT_KERNEL=cta timeout 200 python bench.py --brief --steps 2000 --warmup 200 ) >> $OUT/multiroom_$TAG.log 2>&1
( timeout 600 python bench.py ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
( timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_rollout_cta" -s 8 -c 1 \
    -o $OUT/prof_boss_$TAG -f python bench.py --brief --level BossLevel --envs 32768 --steps 200 --warmup 40 ) > $OUT/ncu_boss_$TAG.log 2>&1
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_rollout_pipe -s 3 -c 1 \
    -o $OUT/prof_rollout_$TAG -f python bench.py --brief --steps 200 --warmup 40 ) > $OUT/ncu_rollout_$TAG.log 2>&1
# launch list of the default bench command (shares only)
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 500 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 200 --warmup 40 --no-cpu-baseline --no-other-configs ) > $OUT/ncu_launches_$TAG.log 2>&1
for tool in memcheck racecheck synccheck; do
  ( timeout 900 compute-sanitizer --tool $tool python scripts/gpu_sanitize.py ) > $OUT/sanitize_${tool}_$TAG.log 2>&1
  echo "exit $?" >> $OUT/sanitize_${tool}_$TAG.log
done
tail -n 3 $OUT/pytest_gpu_$TAG.log
cat $OUT/multiroom_$TAG.log
python - <<PY
import json
d=json.load(open('$OUT/bench_$TAG.json'))
print('value %.4g frac %.4f kernel_frac %.4f e2e %.4g per_step %.4g' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_frac'], d['e2e']['value'], d['per_step_api']['value']))
print(json.dumps(d['other_configs']))
PY
tail -n 5 $OUT/bench_$TAG.err
for tool in memcheck racecheck synccheck; do tail -n 6 $OUT/sanitize_${tool}_$TAG.log; done
