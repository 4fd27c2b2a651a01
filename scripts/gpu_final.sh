#!/bin/bash
# The evidence visit of a round (one B200, run under gpurun):  GPU tests, both bench arms as the driver runs them, the ncu launch
# list of the bench command, one `ncu --set full` capture per hot kernel, compute-sanitizer.   usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi > $OUT/nvidia_smi_$TAG.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
( timeout 900 python bench.py ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
( timeout 600 python bench.py --impl reference ) > $OUT/bench_reference_arm_$TAG.json 2> $OUT/bench_reference_arm_$TAG.err
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > $OUT/smoke_$TAG.log 2>&1
# launch list of the bench command (cold-cache, serialised: compare SHARES)
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 200 --warmup 40 --no-cpu-baseline --no-other-configs ) > $OUT/ncu_launches_$TAG.log 2>&1
# one full capture per hot kernel
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:"^k_rollout$" -s 4 -c 1 \
    -o $OUT/prof_rollout_$TAG -f python bench.py --brief --steps 400 --warmup 40 ) > $OUT/ncu_rollout_$TAG.log 2>&1
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:"^k_gen$|k_rollout_cta" -s 24 -c 3 \
    -o $OUT/prof_boss_$TAG -f python bench.py --brief --level BossLevel --envs 32768 --steps 400 --warmup 40 ) > $OUT/ncu_boss_$TAG.log 2>&1
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_render_rgb" -s 2 -c 1 \
    -o $OUT/prof_rgb_$TAG -f python bench.py --steps 200 --warmup 40 --no-cpu-baseline --no-other-configs ) > $OUT/ncu_rgb_$TAG.log 2>&1
for TOOL in memcheck synccheck; do
  ( timeout 500 compute-sanitizer --tool $TOOL python scripts/gpu_sanitize.py ) > $OUT/sanitize_${TOOL}_$TAG.log 2>&1
done
tail -n 3 $OUT/pytest_gpu_$TAG.log
python - <<PY
import json
d = json.loads([l for l in open('$OUT/bench_$TAG.json') if l.startswith('{')][-1])
print('value %.4g frac %.4f kernel_frac %.4f e2e %.4g per_step %.4g launches %d' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_frac'], d['e2e']['value'], d['per_step_api']['value'], d['gpu_launches']))
print(json.dumps(d.get('other_configs'))[:1500])
r = json.loads([l for l in open('$OUT/bench_reference_arm_$TAG.json') if l.startswith('{')][-1])
print('reference arm: value %.4g %s' % (r['value'], r.get('cpu_baseline')))
PY
tail -n 4 $OUT/bench_$TAG.err $OUT/smoke_$TAG.log $OUT/sanitize_memcheck_$TAG.log $OUT/sanitize_synccheck_$TAG.log
