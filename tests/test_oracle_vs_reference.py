"""Pins the C oracle against the reference itself: /root/reference's unmodified
babyai.levels (levelgen.py, verifier.py, iclr19_levels.py) executing on the
clean-room gym_minigrid shim, compared step by step (obs, reward, done,
direction, mission, full grid, agent pose, carried object, step counters and the
number of RNG draws).  Build-container only."""
import pytest

from common import CONFIG_LEVELS

pytestmark = pytest.mark.reference


@pytest.mark.parametrize('level', CONFIG_LEVELS)
def test_random_policy(level):
    import compare_ref
    for s in range(2):
        compare_ref.compare(level, 4000 + s, 150, 'random', act_seed=s)


@pytest.mark.parametrize('level', ['GoToLocal', 'PickupLoc', 'BossLevel', 'SynthSeq'])
def test_bot_policy(level):
    import compare_ref
    eps = compare_ref.compare(level, 5000, 250, 'bot', act_seed=1)
    assert eps >= 0


def test_reference_level_test_on_shim():
    """A slice of the reference's own test (levelgen.py:496-541) on the shim: determinism per seed."""
    import refenv
    refenv.setup('mt')
    from babyai.levels import level_dict
    for name in CONFIG_LEVELS:
        m0, m1 = level_dict[name](seed=0), level_dict[name](seed=0)
        assert m0.grid == m1.grid and m0.surface == m1.surface
        assert m0.reset()['mission'] == m0.surface


def test_bot_solves_config_levels():
    import selfcheck_bot
    assert selfcheck_bot.run(CONFIG_LEVELS, 3, 'philox')
