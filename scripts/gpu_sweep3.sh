#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_sw.log 2>&1; tail -n 3 $OUT/pytest_gpu_sw.log
for cfg in "0 4" "96 2" "192 2" "192 4" "384 2" "128 1"; do
  set -- $cfg
  echo -n "budget=$1 blocks/SM=$2: "
  BB_GEN_BUDGET=$1 BB_GEN_SMALL_BLOCKS_PER_SM=$2 timeout 300 python bench.py --no-cpu-baseline --steps 4000 --warmup 400 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step %.4f value %.3e errors %d episodes %d' % (d['ms_per_step'], d['value'], d['counters']['errors'], d['counters']['episodes']))
    elif 'rror' in l: print(l.strip()[:200])
"
done 2>&1 | tee $OUT/sweep3.log
