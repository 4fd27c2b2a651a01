#!/bin/bash
# Long rollouts on one B200: no ring may run dry (counters.errors == 0) over 200 000 steps of every level family's supply schedule.
#   usage: bash scripts/gpu_soak.sh <tag>
TAG=${1:-soak}
OUT=gpurun_out
mkdir -p $OUT
for SPEC in "GoToLocal 65536" "PickupLoc 65536" "GoToObjS4 65536" "GoToObjS6 65536" "GoTo 32768" "BossLevel 32768" "Unlock 16384" "KeyCorridorS6R3 16384"; do
  set -- $SPEC
  echo "== $1 $2 envs, 200000 steps" >> $OUT/soak_$TAG.log
  ( timeout 300 python bench.py --brief --envs $2 --steps 200000 --warmup 400 --level $1 ) >> $OUT/soak_$TAG.log 2>&1
done
python - <<PY
import json
for l in open('$OUT/soak_$TAG.log'):
    l = l.strip()
    if l.startswith('=='): print(l, end='  ')
    elif l.startswith('{'):
        d = json.loads(l); c = d['counters']
        print('value %.4e  steps %.3e episodes %d successes %d errors %d' % (d['value'], c['steps'], c['episodes'], c['successes'], c['errors']))
    elif l: print(l[:200])
PY
