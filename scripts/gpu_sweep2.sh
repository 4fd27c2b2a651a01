#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_sw.log 2>&1; tail -n 3 $OUT/pytest_gpu_sw.log
for b in 1 2 4 8; do
  echo -n "gen_small blocks/SM=$b: "
  BB_GEN_SMALL_BLOCKS_PER_SM=$b timeout 300 python bench.py --no-cpu-baseline --steps 2000 --warmup 200 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step %.4f value %.3e k_step %.1f us k_gen %.1f us errors %d' % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms']*1e3, d['roofline']['k_gen_ms']*1e3, d['counters']['errors']))
    elif 'rror' in l: print(l.strip()[:200])
"
done 2>&1 | tee $OUT/sweep2.log
