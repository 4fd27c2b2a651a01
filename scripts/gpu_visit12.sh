#!/bin/bash
# Round-2 twelfth GPU visit: swap-in preload (the next level of an env in its last step is fetched before step_env).
TAG=${1:-r02l}
OUT=gpurun_out
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
run() { echo "== $*" >> $OUT/spec_$TAG.log; ( env "$@" timeout 200 python bench.py --brief --envs 65536 --steps 4000 --warmup 400 --level $LV ) >> $OUT/spec_$TAG.log 2>&1; }
for LV in GoToLocal PickupLoc GoToRedBall GoToObjS6 GoToObjS4; do
  run LV=$LV
  run LV=$LV
done
( timeout 500 ncu --set full --clock-control none --import-source on -k regex:"^k_rollout$" -s 4 -c 2 \
    -o $OUT/prof_rollout_$TAG -f python bench.py --brief --steps 400 --warmup 40 ) > $OUT/ncu_rollout_$TAG.log 2>&1
tail -n 3 $OUT/pytest_gpu_$TAG.log
python - <<PY
import json
for l in open('$OUT/spec_$TAG.log'):
    l=l.strip()
    if l.startswith('=='): print(l, end='  ')
    elif l.startswith('{'):
        d=json.loads(l); print('value %.4e us/step %.3f kernel %.3f frac %.4f errors %d' % (d['value'], d['us_per_step'], d['kernel_us_per_step'], d['roofline_frac'], d['counters']['errors']))
    elif l: print(l[:200])
PY
