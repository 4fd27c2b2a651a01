#!/bin/bash
for cfg in "64 4" "96 4" "128 4" "96 8" "192 8" "64 8"; do
  set -- $cfg
  echo -n "SERIAL budget=$1 blocks/SM=$2: "
  BB_GEN_SERIAL=1 BB_GEN_BUDGET=$1 BB_GEN_SMALL_BLOCKS_PER_SM=$2 timeout 300 python bench.py --no-cpu-baseline --steps 4000 --warmup 400 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step %.4f value %.3e errors %d episodes %d' % (d['ms_per_step'], d['value'], d['counters']['errors'], d['counters']['episodes']))
    elif 'rror' in l: print(l.strip()[:200])
"
done
BB_GEN_SERIAL=1 BB_GEN_BUDGET=96 BB_GEN_SMALL_BLOCKS_PER_SM=4 BB_DEBUG_TIMING=1 python bench.py --no-cpu-baseline --steps 400 --warmup 40 2>&1 | grep "bb timing" | sed -n 4,7p
