"""Glue that runs the reference's OWN, unmodified BabyAI layer
(/root/reference/babyai/levels/*, babyai/rl/utils/penv.py, babyai/evaluate.py,
babyai/bot.py) on top of the clean-room gym / gym_minigrid shim in
oracle/shim/.  This is the strongest parity anchor available:
everything BabyAI-specific that executes is reference code.

TEST INFRASTRUCTURE ONLY.  /root/reference exists only in the build container,
never on the GPU box, so nothing reachable from `-m gpu` tests, smoke() or
bench.py may import this module; it is used by tests/golden/make_golden.py and
by the `not gpu` tests that pin oracle/babyai_oracle.c against the reference.
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, 'shim')
REFERENCE = os.environ.get('BABYAI_REFERENCE', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REFERENCE, 'babyai', 'levels'))


def setup(rng='philox'):
    """Put shim + reference on sys.path, select the RNG back-end, import babyai."""
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE)
    for p in (REFERENCE, SHIM, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gym
    from gym.utils import seeding
    assert os.path.dirname(gym.__file__).startswith(SHIM), 'a different gym is on sys.path'
    seeding.set_backend(rng)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import babyai  # noqa: F401  registers the 105 levels
    return gym


def make_env(level, seed, rng='philox'):
    """gym.make + env.seed(seed), exactly scripts/train_rl.py:57-59."""
    gym = setup(rng)
    env = gym.make('BabyAI-%s-v0' % level)
    # LevelGen.locked_room is only cleared in __init__ (levelgen.py:284); the
    # constructor itself generates one level from an OS-random seed, which can
    # leave a stale value behind.  Start from the documented initial state.
    if hasattr(env, 'locked_room'):
        env.locked_room = None
    env.seed(seed)
    return env
