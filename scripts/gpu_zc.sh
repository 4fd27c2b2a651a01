#!/bin/bash
# final validation + A/B of the zero-copy host path
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for z in 1 2; do echo "zerocopy=$z tests:"; BB_HOST_ZEROCOPY=$z timeout 100 python -m pytest tests -m gpu -x -q -k "host_buffer or smoke or int64" 2>&1 | tail -1; done
for z in 0 1 2; do echo -n "zerocopy=$z: "; BB_HOST_ZEROCOPY=$z timeout 100 python bench.py --no-cpu-baseline --steps 400 --warmup 40 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('e2e %.4e value %.3e errors %d' % (d['e2e']['value'], d['value'], d['counters']['errors']))
"; done
