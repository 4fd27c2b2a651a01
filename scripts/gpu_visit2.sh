#!/bin/bash
# Round-2 second GPU visit: k_gen with one env per ticket, slimmer k_rollout_cta, RGB render kernel.
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
for lv in GoTo BossLevel; do
  echo "== $lv" >> $OUT/multiroom_$TAG.log
  ( timeout 200 python bench.py --brief --level $lv --envs 32768 --steps 2000 --warmup 200 ) >> $OUT/multiroom_$TAG.log 2>&1
done
echo "== GoToLocal BB_ROLLOUT_KERNEL=cta" >> $OUT/multiroom_$TAG.log
( BB_ROLLOUT_KERNEL=cta timeout 200 python bench.py --brief --steps 2000 --warmup 200 ) >> $OUT/multiroom_$TAG.log 2>&1
( timeout 600 python bench.py --no-other-configs ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
( timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_rollout_cta|k_gen<" -s 12 -c 3 \
    -o $OUT/prof_boss_$TAG -f python bench.py --brief --level BossLevel --envs 32768 --steps 200 --warmup 40 ) > $OUT/ncu_boss_$TAG.log 2>&1
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_render_rgb -s 5 -c 1 \
    -o $OUT/prof_rgb_$TAG -f python bench.py --no-other-configs --no-cpu-baseline --steps 200 --warmup 40 ) > $OUT/ncu_rgb_$TAG.log 2>&1
tail -n 3 $OUT/pytest_gpu_$TAG.log
cat $OUT/multiroom_$TAG.log
python - <<PY
import json
d=json.load(open('$OUT/bench_$TAG.json'))
print('value %.4g frac %.4f kernel_frac %.4f e2e %.4g rgb %s' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_frac'], d['e2e']['value'], json.dumps(d['rgb_roofline'])))
PY
tail -n 5 $OUT/bench_$TAG.err
