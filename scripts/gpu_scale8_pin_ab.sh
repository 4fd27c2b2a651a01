#!/bin/bash
# 8-GPU box: is the end-to-end leg better with the ranks pinned to their GPU's NUMA node or not?  Alternating A/B, GoToLocal.
TAG=${1:-r02t}
OUT=gpurun_out
mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 400 --warmup 80 --lean --no-cpu-baseline --no-other-configs; }
for i in 1 2; do
  ( run $((29600 + i)) ) > $OUT/pinab_pin${i}_$TAG.json 2> $OUT/pinab_pin${i}_$TAG.err
  ( BENCH_NO_NUMA_PIN=1 run $((29610 + i)) ) > $OUT/pinab_nopin${i}_$TAG.json 2> $OUT/pinab_nopin${i}_$TAG.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob('$OUT/pinab_*_$TAG.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print('%-32s value %.4g e2e %.4g per-rank %s' % (f.split('/')[-1], d['value'], d['e2e']['value'], json.dumps(d['e2e']['per_rank'])))
    except Exception as ex:
        print(f, 'FAILED', repr(ex)[:200])
PY
