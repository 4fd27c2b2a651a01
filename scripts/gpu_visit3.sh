#!/bin/bash
# Round-2 third GPU visit: SWAR objs_at in step/verify, cached find_reach in connect_all, asynchronous generation passes.
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
for lv in GoTo BossLevel; do
  echo "== $lv" >> $OUT/multiroom_$TAG.log
  ( timeout 200 python bench.py --brief --level $lv --envs 32768 --steps 2000 --warmup 200 ) >> $OUT/multiroom_$TAG.log 2>&1
done
echo "== BossLevel BB_RING_DEPTH=128" >> $OUT/multiroom_$TAG.log
( BB_RING_DEPTH=128 timeout 200 python bench.py --brief --level BossLevel --envs 32768 --steps 2000 --warmup 200 ) >> $OUT/multiroom_$TAG.log 2>&1
echo "== BossLevel BB_GEN_CONCURRENT=0" >> $OUT/multiroom_$TAG.log
( BB_GEN_CONCURRENT=0 timeout 200 python bench.py --brief --level BossLevel --envs 32768 --steps 2000 --warmup 200 ) >> $OUT/multiroom_$TAG.log 2>&1
( timeout 600 python bench.py ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
( timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_rollout_cta|k_gen" -s 14 -c 4 \
    -o $OUT/prof_boss_$TAG -f python bench.py --brief --level BossLevel --envs 32768 --steps 200 --warmup 40 ) > $OUT/ncu_boss_$TAG.log 2>&1
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 3 -c 1 \
    -o $OUT/prof_rollout_$TAG -f python bench.py --brief --steps 200 --warmup 40 ) > $OUT/ncu_rollout_$TAG.log 2>&1
tail -n 3 $OUT/pytest_gpu_$TAG.log
cat $OUT/multiroom_$TAG.log
python - <<PY
import json
d=json.load(open('$OUT/bench_$TAG.json'))
print('value %.4g frac %.4f kernel_frac %.4f e2e %.4g per_step %.4g' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_frac'], d['e2e']['value'], d['per_step_api']['value']))
print(json.dumps(d['other_configs']))
PY
tail -n 5 $OUT/bench_$TAG.err
