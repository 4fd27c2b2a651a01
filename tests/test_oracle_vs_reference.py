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


def _iclr_levels():
    from babyai_b200.levels import ICLR19_LEVELS
    return sorted(ICLR19_LEVELS)


@pytest.mark.parametrize('level', _iclr_levels())
def test_random_policy_every_iclr_level(level):
    """every served ICLR-19 level, 3 seeds x 300 random steps (several episodes on the single-room levels, resets included):
    the CI-length slice of oracle/soak_ref.py (profiles/r02_soak_ref.log has the long run)"""
    import compare_ref
    for s in range(3):
        compare_ref.compare(level, 7100 + 17 * s, 300, 'random', act_seed=10 + s)


@pytest.mark.parametrize('level', _iclr_levels())
def test_bot_policy_every_iclr_level(level):
    """... and the reference bot driving both sides (the success / reward path of every instruction kind)"""
    import compare_ref
    compare_ref.compare(level, 7300, 200, 'bot', act_seed=3)


@pytest.mark.parametrize('level', ['GoToLocal', 'PickupLoc', 'BossLevel', 'SynthSeq', 'GoToImpUnlock', 'Unlock'])
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


def test_reference_place_agent_hang_is_rejected():
    """RoomGrid.place_agent loops `while True` until the cell in front of the agent is empty or a wall
    (gym_minigrid/roomgrid.py place_agent).  MiniBossLevel, seed 698: the 57th level picks a 3x3 room whose two empty
    cells face objects / doors in all four headings, so the reference never returns.  The oracle (and the kernels)
    reject that level instead: same levels up to there, then the oracle carries on."""
    import signal
    import numpy as np
    import oracle as orc
    import refenv
    env = refenv.make_env('MiniBossLevel', 698)
    pool = orc.OraclePool('MiniBossLevel', 1, np.array([698], dtype=np.uint64))
    for k in range(56):
        ob = env.reset()
        oo = pool.reset()
        assert np.array_equal(np.asarray(oo)[0].reshape(7, 7, 3), ob['image']), k
        assert pool.mission(0) == ob['mission']

    class Hung(Exception):
        pass

    def on_alarm(*a):
        raise Hung()
    old = signal.signal(signal.SIGALRM, on_alarm)
    signal.alarm(3)
    try:
        with pytest.raises(Hung):
            env.reset()
    finally:
        signal.alarm(0)
        signal.signal(signal.SIGALRM, old)
    # the room the reference is stuck in: no empty cell with an empty / wall cell in front of it
    e = env.unwrapped
    stuck = []
    for j in range(e.num_rows):
        for i in range(e.num_cols):
            room = e.get_room(i, j)
            ok = False
            for y in range(room.top[1] + 1, room.top[1] + room.size[1] - 1):
                for x in range(room.top[0] + 1, room.top[0] + room.size[0] - 1):
                    if e.grid.get(x, y) is not None:
                        continue
                    for dx, dy in ((1, 0), (0, 1), (-1, 0), (0, -1)):
                        c = e.grid.get(x + dx, y + dy)
                        ok = ok or c is None or c.type == 'wall'
            if not ok:
                stuck.append((i, j))
    assert stuck, 'expected a room without any admissible agent pose'
    pool.reset()        # returns: the oracle rejected the unsatisfiable level
    assert pool.state(0)[1]['attempts'] >= 58
