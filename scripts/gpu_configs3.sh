#!/bin/bash
run() { echo -n "$1 envs=$2 D=$BB_RING_DEPTH G=$BB_GEN_PERIOD: "; timeout 900 python bench.py --no-cpu-baseline --level $1 --envs $2 --steps $3 --warmup $4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.3e ms_per_step %.4f per_step_api %.3e errors %d episodes %d' % (d['value'], d['ms_per_step'], d['per_step_api']['value'], d['counters']['errors'], d['counters']['episodes']))
    elif 'rror' in l: print(l.strip()[:300])
"; }
for cfg in "32 8" "64 16" "64 32" "128 32"; do set -- $cfg; export BB_RING_DEPTH=$1 BB_GEN_PERIOD=$2; run BossLevel 32768 1600 160; run GoTo 32768 1600 160; done
