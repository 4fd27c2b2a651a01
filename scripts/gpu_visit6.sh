#!/bin/bash
# Round-2 sixth GPU visit: lane-per-level k_gen, 48-register k_rollout_cta, ring depth 512.
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
run() { echo "== $*" >> $OUT/multiroom_$TAG.log; ( env "$@" timeout 200 python bench.py --brief --envs 32768 --steps 2000 --warmup 200 --level $LV ) >> $OUT/multiroom_$TAG.log 2>&1; }
for LV in GoTo BossLevel; do
  run LV=$LV
  run LV=$LV BB_GEN_LANES=4
  run LV=$LV BB_GEN_LANES=16
  run LV=$LV BB_GEN_LANES=32
  run LV=$LV BB_RING_DEPTH=256
  run LV=$LV BB_GEN_BESIDE_BLOCKS_PER_SM=1
  run LV=$LV BB_GEN_BESIDE_BLOCKS_PER_SM=4
  run LV=$LV BB_GEN_CONCURRENT=0
done
( timeout 600 python bench.py ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
( timeout 500 ncu --set full --clock-control none --import-source on -k regex:"^k_gen$|k_rollout_cta" -s 20 -c 4 \
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
d=json.load(open('$OUT/bench_$TAG.json'))
print('value %.4g frac %.4f kernel_frac %.4f e2e %.4g per_step %.4g' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_frac'], d['e2e']['value'], d['per_step_api']['value']))
print(json.dumps(d['other_configs'])[:1800])
PY
tail -n 5 $OUT/bench_$TAG.err
