#!/bin/bash
# 8-GPU box: BASELINE config 5 as specified (BossLevel, 262 144 envs = 8 x 32 768, NCCL counter all-gather, scaling 1/2/4/8)
# and the bench workload at N = 4 / 8 (NUMA-pinned ranks: the end-to-end leg must scale).  usage: bash scripts/gpu_scale8.sh [tag]
TAG=${1:-r02s}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo8_$TAG.txt 2>&1
run() {  # N, extra args...
  N=$1; shift
  if [ "$N" = 1 ]; then python bench.py --gpus 1 "$@"
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N "$@"; fi
}
for N in 8 4 2 1; do
  ( run $N --level BossLevel --envs 32768 --steps 2000 --warmup 200 --lean --no-cpu-baseline --no-other-configs ) > $OUT/scale_boss_${N}_$TAG.json 2> $OUT/scale_boss_${N}_$TAG.err
done
for N in 8 4; do
  ( run $N --steps 2000 --warmup 200 --lean --no-cpu-baseline --no-other-configs ) > $OUT/scale_gotolocal_${N}_$TAG.json 2> $OUT/scale_gotolocal_${N}_$TAG.err
done
( BENCH_NO_NUMA_PIN=1 run 8 --steps 2000 --warmup 200 --lean --no-cpu-baseline --no-other-configs ) > $OUT/scale_gotolocal_8nopin_$TAG.json 2> $OUT/scale_gotolocal_8nopin_$TAG.err
python - <<PY
import json, glob
for f in sorted(glob.glob('$OUT/scale_*_$TAG.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print('%-40s n_gpus %d value %.4g ms/step %.5f e2e %.4g per-rank e2e %s numa %s counters %s' % (f.split('/')[-1], d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], json.dumps(d['e2e']['per_rank']), json.dumps(d['config'].get('numa')), json.dumps(d['counters'])))
    except Exception as ex:
        print(f, 'FAILED', repr(ex)[:200])
PY
tail -n 3 $OUT/scale_*_$TAG.err | tail -n 30
