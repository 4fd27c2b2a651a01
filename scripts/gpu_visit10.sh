#!/bin/bash
# Round-2 tenth GPU visit: k_rollout_cta v2 (stepper runs ahead, producer/consumer barriers) vs v1; k_gen with GenCtx in registers.
TAG=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
run() { echo "== $*" >> $OUT/multiroom_$TAG.log; ( env "$@" timeout 200 python bench.py --brief --envs 32768 --steps 2000 --warmup 200 --level $LV ) >> $OUT/multiroom_$TAG.log 2>&1; }
for LV in GoTo BossLevel; do
  run LV=$LV
  run LV=$LV BB_ROLLOUT_KERNEL=cta1
  run LV=$LV BB_GEN_CONCURRENT=0
  run LV=$LV BB_GEN_CONCURRENT=0 BB_ROLLOUT_KERNEL=cta1
  run LV=$LV BB_GEN_CONCURRENT=0 BB_GEN_BLOCKS_PER_SM=16
  run LV=$LV BB_GEN_BESIDE_BLOCKS_PER_SM=4
  run LV=$LV BB_GEN_BESIDE_BLOCKS_PER_SM=8
  run LV=$LV BB_GEN_BESIDE_BLOCKS_PER_SM=8 BB_GEN_CHAIN_CAP=0
done
( timeout 500 ncu --set full --clock-control none --import-source on -k regex:"^k_gen$|k_rollout_cta" -s 24 -c 4 \
    -o $OUT/prof_boss_$TAG -f python bench.py --brief --level BossLevel --envs 32768 --steps 400 --warmup 40 ) > $OUT/ncu_boss_$TAG.log 2>&1
tail -n 3 $OUT/pytest_gpu_$TAG.log
python - <<PY
import json
for l in open('$OUT/multiroom_$TAG.log'):
    l=l.strip()
    if l.startswith('=='): print(l, end='  ')
    elif l.startswith('{'):
        d=json.loads(l); print('value %.3e us/step %.2f kernel %.2f refill/launch %.3f ms errors %d' % (d['value'], d['us_per_step'], d['kernel_us_per_step'], d['refill_ms_per_launch'], d['counters']['errors']))
    elif l: print(l[:200])
PY
