#!/bin/bash
run() { timeout 300 python bench.py --no-cpu-baseline --steps 4000 --warmup 400 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step %.4f value %.3e rollout %.1f us refill/launch %.1f us errors %d' % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms']*1e3, d['roofline']['refill_ms_per_launch']*1e3, d['counters']['errors']))
    elif 'rror' in l: print(l.strip()[:200])
"; }
echo -n "default: "; run
echo -n "concurrent: "; BB_GEN_CONCURRENT=1 run
echo -n "concurrent blocks/SM=2 budget 256: "; BB_GEN_CONCURRENT=1 BB_GEN_SMALL_BLOCKS_PER_SM=2 BB_GEN_BUDGET=256 run
echo -n "R=3 D=176 budget 192: "; BB_RING_DEPTH=176 BB_REFILL_EVERY=3 BB_GEN_BUDGET=192 run
echo -n "blocks/SM=8: "; BB_GEN_SMALL_BLOCKS_PER_SM=8 run
