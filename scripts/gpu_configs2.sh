#!/bin/bash
( timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -n 2
run() { echo -n "$1 envs=$2: "; timeout 900 python bench.py --no-cpu-baseline --level $1 --envs $2 --steps $3 --warmup $4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.3e ms_per_step %.4f per_step_api %.3e errors %d episodes %d successes %d' % (d['value'], d['ms_per_step'], d['per_step_api']['value'], d['counters']['errors'], d['counters']['episodes'], d['counters']['successes']))
    elif 'rror' in l: print(l.strip()[:300])
"; }
run GoTo 32768 800 80
run BossLevel 32768 800 80
