#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_ab.log 2>&1; tail -n 3 $OUT/pytest_gpu_ab.log
for k in cols lane staged; do
  echo "== BB_STEP_KERNEL=$k"
  BB_STEP_KERNEL=$k timeout 300 python bench.py --no-cpu-baseline --steps 2000 --warmup 200 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step %.4f value %.3e k_step %.1f us k_gen %.1f us e2e %.3e errors %d' % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms']*1e3, d['roofline']['k_gen_ms']*1e3, d['e2e']['value'], d['counters']['errors']))
    elif 'rror' in l: print(l.strip()[:200])
"
done 2>&1 | tee $OUT/ab.log
