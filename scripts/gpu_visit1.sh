#!/bin/bash
# Round-2 first GPU visit: parity of the refactored kernels, the new bench line, counter-backed "before" captures of the
# multi-room kernels at BASELINE config 5's per-GPU size (VERDICT r1 missing #6), and a source-level capture of k_rollout.
# usage (repo root, on the GPU box): bash scripts/gpu_visit1.sh [tag]
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi > $OUT/nvsmi_$TAG.txt 2>&1
lscpu | head -25 > $OUT/lscpu_$TAG.txt 2>&1
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke_$TAG.log 2>&1
echo "smoke exit $?" >> $OUT/smoke_$TAG.log
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
( timeout 600 python bench.py ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
# multi-room "before" captures: BossLevel, 32 768 envs (round-1 path: lane-per-env k_rollout on 22x22 staging + k_gen beside it)
( BB_ROLLOUT_KERNEL=lane timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_rollout|k_gen" -s 10 -c 3 \
    -o $OUT/prof_boss_lane_$TAG -f python bench.py --brief --level BossLevel --envs 32768 --steps 200 --warmup 40 ) > $OUT/ncu_boss_lane_$TAG.log 2>&1
# ... and the new k_rollout_cta
( timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_rollout_cta" -s 6 -c 1 \
    -o $OUT/prof_boss_cta_$TAG -f python bench.py --brief --level BossLevel --envs 32768 --steps 200 --warmup 40 ) > $OUT/ncu_boss_cta_$TAG.log 2>&1
# single-room: source-level capture of the fused k_rollout (stall reasons per line)
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 3 -c 1 \
    -o $OUT/prof_rollout_$TAG -f python bench.py --brief --steps 200 --warmup 40 ) > $OUT/ncu_rollout_$TAG.log 2>&1
for lv in GoTo BossLevel; do for k in lane cta; do
  echo "== $lv BB_ROLLOUT_KERNEL=$k" >> $OUT/multiroom_$TAG.log
  ( BB_ROLLOUT_KERNEL=$k timeout 200 python bench.py --brief --level $lv --envs 32768 --steps 2000 --warmup 200 ) >> $OUT/multiroom_$TAG.log 2>&1
done; done
echo "== GoToLocal BB_ROLLOUT_KERNEL=cta (single-room through the CTA kernel, refill passes)" >> $OUT/multiroom_$TAG.log
( BB_ROLLOUT_KERNEL=cta timeout 200 python bench.py --brief --steps 2000 --warmup 200 ) >> $OUT/multiroom_$TAG.log 2>&1
tail -n 3 $OUT/smoke_$TAG.log $OUT/pytest_gpu_$TAG.log
cat $OUT/bench_$TAG.json
tail -n 5 $OUT/bench_$TAG.err
cat $OUT/multiroom_$TAG.log
ls -la $OUT
