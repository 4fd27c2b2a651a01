#!/bin/bash
# fused generator warp inside k_rollout vs refill passes
run() {
  timeout 300 python bench.py --no-cpu-baseline --steps 4000 --warmup 400 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step %.4f value %.3e rollout %.1f us refill/launch %.1f us per-step-api %.3e e2e %.3e errors %d' % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms']*1e3, d['roofline']['refill_ms_per_launch']*1e3, d['per_step_api']['value'], d['e2e']['value'], d['counters']['errors']))
    elif 'rror' in l: print(l.strip()[:200])
"
}
echo -n "fused default (budget 8, min_active 16): "; run
echo -n "fused min_active 8: "; BB_GEN_MIN_ACTIVE=8 run
echo -n "fused min_active 24: "; BB_GEN_MIN_ACTIVE=24 run
echo -n "fused budget 2: "; BB_GEN_BUDGET=2 run
echo -n "fused budget 3 min_active 1: "; BB_GEN_BUDGET=3 BB_GEN_MIN_ACTIVE=1 run
echo -n "not fused (refill passes every 2nd launch): "; BB_GEN_FUSED=0 run
for lv in PickupLoc GoToRedBall GoToObjS4 GoToLocalS5N2; do echo -n "fused $lv: "; timeout 300 python bench.py --no-cpu-baseline --steps 2000 --warmup 200 --level $lv 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step %.4f value %.3e rollout %.1f us errors %d episodes %d' % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms']*1e3, d['counters']['errors'], d['counters']['episodes']))
    elif 'rror' in l: print(l.strip()[:200])
"; done
