#!/bin/bash
run() { timeout 900 python bench.py --no-cpu-baseline --level $1 --envs 32768 --steps 1600 --warmup 160 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.3e ms_per_step %.4f rollout-kernel %.1f us refill %.1f us errors %d episodes %d' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']*1e3, d['roofline']['refill_ms_per_launch']*1e3, d['counters']['errors'], d['counters']['episodes']))
    elif 'rror' in l: print(l.strip()[:300])
"; }
export BB_PERSIST_MAX_CELLS=1152
for lv in BossLevel GoTo; do
  echo -n "$lv D=128 R=2: "; BB_RING_DEPTH=128 BB_REFILL_EVERY=2 run $lv
  echo -n "$lv D=96 R=1: "; BB_RING_DEPTH=96 BB_REFILL_EVERY=1 run $lv
done
