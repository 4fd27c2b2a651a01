#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_sw.log 2>&1; tail -n 3 $OUT/pytest_gpu_sw.log
for cfg in "64 4" "48 4" "96 2" "64 2" "128 1"; do
  set -- $cfg
  echo -n "budget=$1 blocks/SM=$2: "
  BB_GEN_BUDGET=$1 BB_GEN_SMALL_BLOCKS_PER_SM=$2 timeout 300 python bench.py --no-cpu-baseline --steps 4000 --warmup 400 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step %.4f value %.3e rollout %.1f us refill %.1f us e2e %.3e errors %d' % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms']*1e3, d['roofline']['refill_ms_per_launch']*1e3, d['e2e']['value'], d['counters']['errors']))
    elif 'rror' in l: print(l.strip()[:200])
"
done 2>&1 | tee $OUT/sweep5.log
