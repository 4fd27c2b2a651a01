"""Long fuzz of the kernels' per-environment source (host build, tests/hostemu) against the C oracle over every served
level: random and interaction-heavy action mixes, full hidden state compared every few steps, and deep walks down each
env's level stream (consecutive resets).  TEST INFRASTRUCTURE (CPU only); the short versions live in tests/test_hostemu.py.

    python scripts/fuzz_hostemu.py steps  <n_envs> <steps> <seed_base> [levels]
    python scripts/fuzz_hostemu.py resets <n_envs> <resets> <seed_base> [levels]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'hostemu')):
    sys.path.insert(0, p)
import hostemu  # noqa: E402
import oracle as orc  # noqa: E402
from babyai_b200.levels import ICLR19_LEVELS as LEVELS, detokenize, level_spec  # noqa: E402  (the levels the C oracle covers)
from common import compare_pools  # noqa: E402

MIXES = [None, [0.12, 0.12, 0.30, 0.17, 0.14, 0.13, 0.02], [0.05, 0.05, 0.45, 0.15, 0.15, 0.15, 0.0]]


def main():
    mode, n, count, base = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    levels = sys.argv[5].split(',') if len(sys.argv) > 5 else sorted(LEVELS)
    bad = []
    for level in levels:
        t0 = time.time()
        try:
            if mode == 'steps':
                for k, p in enumerate(MIXES):
                    seeds = np.arange(n, dtype=np.uint64) * 13 + base + 1000 * k
                    compare_pools(orc.OraclePool(level, n, seeds), hostemu.HostEmuPool(level_spec(level), n, seeds), n, count,
                                  act_seed=base + k, action_p=p, mission_a=lambda q, i: q.mission(i),
                                  mission_b=lambda q, i: detokenize(q.tokens(i)))
            else:
                seeds = np.arange(n, dtype=np.uint64) * 7 + base
                e, o = hostemu.HostEmuPool(level_spec(level), n, seeds), orc.OraclePool(level, n, seeds)
                for k in range(count):
                    assert np.array_equal(e.reset(), np.asarray(o.reset())), k
                    if k % 50 == 0:
                        assert all(detokenize(e.tokens(i)) == o.mission(i) for i in range(n)), k
        except AssertionError as ex:
            bad.append((level, str(ex)[:200]))
            print('%20s MISMATCH %s' % (level, str(ex)[:200]), flush=True)
            continue
        print('%20s ok  %.0f s' % (level, time.time() - t0), flush=True)
    print('mismatches:', bad)
    sys.exit(1 if bad else 0)


if __name__ == '__main__':
    main()
