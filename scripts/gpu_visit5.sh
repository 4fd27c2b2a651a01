#!/bin/bash
# Round-2 fifth GPU visit: bonus levels + verifier modes on the GPU, sanitizers after the barrier changes.
TAG=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
for tool in memcheck racecheck synccheck; do
  ( timeout 900 compute-sanitizer --tool $tool python scripts/gpu_sanitize.py ) > $OUT/sanitize_${tool}_$TAG.log 2>&1
  echo "exit $?" >> $OUT/sanitize_${tool}_$TAG.log
done
( timeout 600 python bench.py ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -n 6 $OUT/pytest_gpu_$TAG.log
for tool in memcheck racecheck synccheck; do grep -E "SUMMARY|driver:|exit" $OUT/sanitize_${tool}_$TAG.log | tail -n 3; done
python - <<PY
import json
d=json.load(open('$OUT/bench_$TAG.json'))
print('value %.4g frac %.4f kernel_frac %.4f e2e %.4g per_step %.4g' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_frac'], d['e2e']['value'], d['per_step_api']['value']))
print(json.dumps(d['other_configs'])[:1500])
PY
