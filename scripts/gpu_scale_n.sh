#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for k in cols lane; do for n in 8192 32768 65536 131072 262144; do
  echo -n "kernel=$k n=$n: "
  BB_STEP_KERNEL=$k timeout 300 python bench.py --no-cpu-baseline --steps 800 --warmup 80 --envs $n 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step %.4f value %.3e k_step %.1f us k_gen %.1f us errors %d' % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms']*1e3, d['roofline']['k_gen_ms']*1e3, d['counters']['errors']))
    elif 'rror' in l: print(l.strip()[:200])
"
done; done 2>&1 | tee $OUT/scale_n.log
