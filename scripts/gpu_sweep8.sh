#!/bin/bash
# k_gen_small: refill period, round budget, sparse-warp stop threshold, blocks per SM
for cfg in "2 8 0 4" "2 8 16 4" "1 8 16 4" "1 8 24 4" "1 8 8 4" "1 8 16 3" "1 8 16 6" "1 2 16 4" "1 8 28 4"; do
  set -- $cfg
  echo -n "refill_every=$1 budget=$2 min_active=$3 blocks/SM=$4: "
  BB_REFILL_EVERY=$1 BB_GEN_BUDGET=$2 BB_GEN_MIN_ACTIVE=$3 BB_GEN_SMALL_BLOCKS_PER_SM=$4 timeout 300 python bench.py --no-cpu-baseline --steps 4000 --warmup 400 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step %.4f value %.3e rollout %.1f us refill/launch %.1f us per-step-api %.3e errors %d' % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms']*1e3, d['roofline']['refill_ms_per_launch']*1e3, d['per_step_api']['value'], d['counters']['errors']))
    elif 'rror' in l: print(l.strip()[:200])
"
done
