#!/bin/bash
# k_gen_small (round-based) tuning: ring depth, refill period, round budget, blocks per SM
for cfg in "128 2 8 4" "128 2 8 2" "128 2 8 3" "128 2 8 6" "128 1 8 4" "128 2 4 4" "128 2 16 4" "128 1 4 2" "192 3 8 4"; do
  set -- $cfg
  echo -n "D=$1 refill_every=$2 budget=$3 blocks/SM=$4: "
  BB_RING_DEPTH=$1 BB_REFILL_EVERY=$2 BB_GEN_BUDGET=$3 BB_GEN_SMALL_BLOCKS_PER_SM=$4 timeout 300 python bench.py --no-cpu-baseline --steps 4000 --warmup 400 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step %.4f value %.3e rollout %.1f us refill/launch %.1f us per-step-api %.3e errors %d' % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms']*1e3, d['roofline']['refill_ms_per_launch']*1e3, d['per_step_api']['value'], d['counters']['errors']))
    elif 'rror' in l: print(l.strip()[:200])
"
done
