#!/bin/bash
# First GPU visit of round 2 (about 12 GPU-minutes): everything written after round 1's GPU budget was spent runs here
# for the first time -- tests/test_zz_gpu_widening.py (28 more golden replays, GoToImpUnlock / Unlock against the oracle,
# DeviceParallelEnv + ObssPreprocessor), bench.py's learner_path leg -- plus the throughput of the two new levels.
# usage (repo root, on the GPU box): bash scripts/gpu_round2_first.sh [tag]
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi > $OUT/nvsmi_$TAG.txt 2>&1
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke_$TAG.log 2>&1
echo "smoke exit $?" >> $OUT/smoke_$TAG.log
( timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu_$TAG.log 2>&1      # no -x: see every failure of the new file
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
( timeout 600 python bench.py ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
( timeout 600 python bench.py --impl reference --steps 200 --warmup 10 ) > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
for cfg in "GoToImpUnlock 32768" "Unlock 32768" "BossLevel 32768" "GoTo 32768"; do set -- $cfg
  echo -n "$1 envs=$2: " >> $OUT/configs_$TAG.log
  timeout 600 python bench.py --no-cpu-baseline --no-probe --level $1 --envs $2 --steps 1600 --warmup 160 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.3e per_step_api %.3e e2e %.3e learner %s errors %d episodes %d' % (d['value'], d['per_step_api']['value'], d['e2e']['value'], d.get('learner_path'), d['counters']['errors'], d['counters']['episodes']))
    elif 'rror' in l: print(l.strip()[:300])
" >> $OUT/configs_$TAG.log
done
# k_rollout2 (two lanes per environment, BB_ROLLOUT_LANES=2): never run before -- correctness first, then the A/B
( BB_TEST_ROLLOUT2=1 timeout 900 python -m pytest tests/test_zz_gpu_widening.py -q -k rollout2 ) > $OUT/pytest_rollout2_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_rollout2_$TAG.log
for lanes in 1 2; do for lv in GoToLocal PickupLoc BossLevel; do
  echo -n "$lv lanes=$lanes: " >> $OUT/ab_lanes_$TAG.log
  BB_ROLLOUT_LANES=$lanes timeout 300 python bench.py --no-cpu-baseline --no-probe --steps 2000 --warmup 200 --level $lv $( [ $lv = BossLevel ] && echo --envs 32768 ) 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.3e rollout kernel %.1f us frac %.3f errors %d' % (d['value'], d['roofline']['kernel_ms']*1e3, d['roofline']['frac'], d['counters']['errors']))
" >> $OUT/ab_lanes_$TAG.log
done; done
cat $OUT/ab_lanes_$TAG.log; tail -n 5 $OUT/pytest_rollout2_$TAG.log
cat $OUT/configs_$TAG.log
tail -n 15 $OUT/pytest_gpu_$TAG.log; tail -n 2 $OUT/smoke_$TAG.log
cat $OUT/bench_$TAG.json; tail -n 5 $OUT/bench_$TAG.err
