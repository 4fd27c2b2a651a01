"""Long-running comparison of the C oracle with the reference's own BabyAI layer (unmodified babyai.levels on the
gym_minigrid shim, Philox back-end) over every served level.  TEST INFRASTRUCTURE ONLY; build container only
(needs /root/reference).  The short, per-commit version of these checks is tests/test_oracle_vs_reference.py.

    python oracle/soak_ref.py random 60 1000 110000     # 60 seeds x 1000 uniformly random actions per level
    python oracle/soak_ref.py bot    25  600 210000     # reference bot (25 % random perturbation)
    python oracle/soak_ref.py reset  12  150 310000     # 150 consecutive resets per seed: deep into each level stream

Every step (or reset) compares observation, direction, mission, reward, done, every grid cell, agent pose, carried
object, step counters and the number of RNG draws (compare_ref.py).  A reference call that does not return within the
alarm is reported as REF HANG (RoomGrid.place_agent's unbounded loop); the oracle must return there.

Last full run (round 1): random 45 x 60 x 1000, bot 45 x 25 x 600, reset 45 x 12 x 150 -- 0 mismatches; one REF HANG
(MiniBossLevel seed 310117, 81st level) next to the one pinned in the test-suite (seed 698, 57th level).
"""
import os
import signal
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import compare_ref  # noqa: E402
import oracle as orc  # noqa: E402
import refenv  # noqa: E402
from babyai_b200.levels import ICLR19_LEVELS, LEVELS as _ALL  # noqa: E402  (the table of served level ids; no device code is touched)
LEVELS = {k: _ALL[k] for k in ICLR19_LEVELS}      # what the C oracle covers; the bonus levels: tests/test_bonus_levels.py


class Hung(Exception):
    pass


def _on_alarm(*a):
    raise Hung()


def soak_steps(policy, nseeds, steps, base):
    bad = []
    for level in sorted(LEVELS):
        for s in range(nseeds):
            signal.alarm(120)
            try:
                compare_ref.compare(level, base + 17 * s, steps, policy, act_seed=s)
            except Hung:
                print('REF HANG', level, base + 17 * s, flush=True)
            except AssertionError as e:
                print('MISMATCH', level, base + 17 * s, str(e)[:300], flush=True)
                bad.append((level, base + 17 * s))
            finally:
                signal.alarm(0)
        print(level, 'done', flush=True)
    return bad


def soak_resets(nseeds, resets, base):
    bad = []
    for level in sorted(LEVELS):
        for s in range(nseeds):
            seed = base + 13 * s
            env = refenv.make_env(level, seed, 'philox')
            pool = orc.OraclePool(level, 1, seeds=[seed])
            for k in range(resets):
                signal.alarm(20)
                try:
                    obs = env.reset()
                except Hung:
                    print('REF HANG', level, seed, k, flush=True)
                    signal.alarm(0)
                    pool.reset()            # the oracle rejects the unsatisfiable level and carries on
                    break
                finally:
                    signal.alarm(0)
                o = pool.reset()
                try:
                    assert np.array_equal(obs['image'], o[0]), 'obs'
                    assert obs['mission'] == pool.mission(0), 'mission'
                    compare_ref.check_state(env.unwrapped, pool, (level, seed, k))
                except AssertionError as e:
                    print('MISMATCH', level, seed, k, str(e)[:200], flush=True)
                    bad.append((level, seed, k))
                    break
        print(level, 'done', flush=True)
    return bad


if __name__ == '__main__':
    mode, nseeds, count, base = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    if len(sys.argv) > 5:                                  # optional: comma-separated subset of the served levels
        LEVELS = {k: LEVELS[k] for k in sys.argv[5].split(',')}
    signal.signal(signal.SIGALRM, _on_alarm)
    t0 = time.time()
    bad = soak_resets(nseeds, count, base) if mode == 'reset' else soak_steps(mode, nseeds, count, base)
    print('mismatches:', bad, 'seconds:', round(time.time() - t0))
    sys.exit(1 if bad else 0)
