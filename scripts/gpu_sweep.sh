#!/bin/bash
# ring depth / generation period sweep (bench only)
OUT=gpurun_out; mkdir -p $OUT
for cfg in "32 4" "32 8" "32 16" "64 8" "64 16" "64 32"; do
  set -- $cfg
  echo "== D=$1 G=$2"
  BB_RING_DEPTH=$1 BB_GEN_PERIOD=$2 timeout 300 python bench.py --no-cpu-baseline --steps 2000 --warmup 200 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step %.4f value %.3e k_step %.1f us k_gen %.1f us errors %d' % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms']*1e3, d['roofline']['k_gen_ms']*1e3, d['counters']['errors']))
    elif 'rror' in l: print(l.strip()[:200])
"
done 2>&1 | tee $OUT/sweep.log
