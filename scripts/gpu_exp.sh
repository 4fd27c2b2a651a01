#!/bin/bash
run() { timeout 300 python bench.py --no-cpu-baseline "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step %.4f value %.3e errors %d' % (d['ms_per_step'], d['value'], d['counters']['errors']))
    elif 'rror' in l: print(l.strip()[:200])
"; }
echo -n "no refill (k_rollout alone), 320 steps: "; BB_DEBUG_NO_REFILL=1 run --steps 320 --warmup 40
echo -n "with refill, 320 steps: "; run --steps 320 --warmup 40
echo -n "with refill, 4000 steps: "; run --steps 4000 --warmup 400
for n in 16384 32768 131072; do echo -n "envs=$n: "; run --steps 2000 --warmup 200 --envs $n; done
