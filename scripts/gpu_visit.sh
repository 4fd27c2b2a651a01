#!/bin/bash
# One single-GPU measurement visit (run under gpurun):  GPU tests, then `bench.py --brief` for every (level, variant) pair.
#   usage: bash scripts/gpu_visit.sh <tag> <envs> "<level> <level> ..." "<VAR=val VAR=val>" ["<variant>" ...]
#   e.g.   bash scripts/gpu_visit.sh r02j 32768 "GoTo BossLevel" "" "BB_GEN_CONCURRENT=0" "BB_GEN_BESIDE_BLOCKS_PER_SM=4"
# A variant is a list of environment settings (the pool's tuning knobs); "" = the defaults.  Output:
# gpurun_out/pytest_gpu_<tag>.log, gpurun_out/sweep_<tag>.log (one JSON line per run) and a table on stdout.
# (The round-2 visits r02a .. r02l were run with per-visit copies of this script; git history has them as scripts/gpu_visit<N>.sh.)
TAG=$1; ENVS=$2; LEVELS=$3; shift 3
OUT=gpurun_out
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
[ $# -eq 0 ] && set -- ""
for LV in $LEVELS; do
  for V in "$@"; do
    echo "== LV=$LV $V" >> $OUT/sweep_$TAG.log
    ( env $V timeout 300 python bench.py --brief --envs $ENVS --steps 2000 --warmup 200 --level $LV ) >> $OUT/sweep_$TAG.log 2>&1
  done
done
tail -n 3 $OUT/pytest_gpu_$TAG.log
python - <<PY
import json
for l in open('$OUT/sweep_$TAG.log'):
    l = l.strip()
    if l.startswith('=='): print(l, end='  ')
    elif l.startswith('{'):
        d = json.loads(l)
        print('value %.4e us/step %.3f kernel %.3f refill/launch %.3f ms frac %.4f errors %d' % (d['value'], d['us_per_step'], d['kernel_us_per_step'], d['refill_ms_per_launch'], d['roofline_frac'], d['counters']['errors']))
    elif l: print(l[:200])
PY
