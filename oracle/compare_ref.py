"""Step-by-step comparison of the C oracle (oracle/babyai_oracle.c) with the
reference's own BabyAI layer running on the gym_minigrid shim, Philox back-end.
TEST INFRASTRUCTURE ONLY; needs /root/reference (build container only).

Checked at reset and after every step: observation image bytes, direction,
mission string, reward (as float32), done, and the full hidden state: every
grid cell's (type, color, state), agent pose, carried object, step_count,
max_steps and the number of RNG draws consumed.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refenv  # noqa: E402
import oracle as orc  # noqa: E402


def ref_state(env):
    g = env.grid
    img = g.encode()                       # [x, y, 3]
    packed = (img[:, :, 0] | (img[:, :, 1] << 3) | (img[:, :, 2] << 6)).astype(np.uint8).T   # [y, x]
    car = 0
    if env.carrying is not None:
        t = env.carrying.encode()
        car = t[0] | (t[1] << 3) | (t[2] << 6)
    return packed, dict(agent_x=int(env.agent_pos[0]), agent_y=int(env.agent_pos[1]), agent_dir=int(env.agent_dir),
                        carrying=int(car), step_count=int(env.step_count), max_steps=int(env.max_steps),
                        draws=int(env.np_random.draws) & 0x7FFFFFFF)


def check_state(env, pool, tag, check_draws=True):
    g0, i0 = ref_state(env)
    g1, i1 = pool.state(0)
    i1 = dict(i1)
    i1.pop('attempts')
    if not check_draws:                    # (a pool that generates levels ahead has consumed more draws)
        i0.pop('draws'); i1.pop('draws')
    assert np.array_equal(g0, g1), (tag, 'grid', g0, g1)
    assert i0 == i1, (tag, i0, i1)


def compare(level, seed, steps, policy='random', act_seed=0, verbose=False, make_pool=None, state_at_reset=True, check_draws=True):
    """policy: 'random' | 'bot' (reference bot, 25% random perturbation).  make_pool(level, seed) -> a one-env pool with the
    oracle's interface (default: the C oracle); the bonus levels are checked with the host build of the kernel source."""
    env = refenv.make_env(level, seed, 'philox')
    pool = orc.OraclePool(level, 1, seeds=[seed]) if make_pool is None else make_pool(level, seed)
    rng = np.random.RandomState(act_seed)
    obs = env.reset()
    o = pool.reset()
    assert np.array_equal(obs['image'], o[0]), (level, seed, 'reset obs')
    assert obs['mission'] == pool.mission(0), (obs['mission'], pool.mission(0))
    if state_at_reset:
        check_state(env, pool, (level, seed, 'reset'), check_draws)
    bot = None
    if policy == 'bot':
        from babyai.bot import Bot
        bot = Bot(env)
    episodes = 0
    last = None
    for t in range(steps):
        if bot is not None and rng.rand() < 0.75:
            try:
                a = int(bot.replan(last))
            except Exception:
                a = int(rng.randint(0, 7))
                bot = None
        else:
            a = int(rng.randint(0, 7))
            if bot is not None:
                # the bot cannot recover from opened boxes / closed doors (eval_bot.py:131-139)
                fc = env.grid.get(*env.front_pos)
                if a == 5 and fc is not None and (fc.type == 'box' or (fc.type == 'door' and fc.is_open)):
                    a = 6
        last = a
        obs, reward, done, _ = env.step(a)
        at_reset = False
        if done:
            obs = env.reset()                    # penv.py:9-10
            episodes += 1
            at_reset = True
            if policy == 'bot':
                from babyai.bot import Bot
                bot = Bot(env)
                last = None
        o, r, d = pool.step([a])
        tag = (level, seed, t, a)
        assert bool(d[0]) == bool(done), (tag, 'done', d[0], done)
        assert np.float32(reward) == r[0], (tag, 'reward', reward, r[0])
        assert np.array_equal(obs['image'], o[0]), (tag, 'obs')
        assert obs['direction'] == pool.direction[0], (tag, 'dir')
        assert obs['mission'] == pool.mission(0), (tag, obs['mission'], pool.mission(0))
        if state_at_reset or not at_reset:
            check_state(env, pool, tag, check_draws)
    return episodes


if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--levels', default='GoToRedBall,GoToLocal,PickupLoc,GoTo,BossLevel')
    ap.add_argument('--seeds', type=int, default=5)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--policy', default='random')
    a = ap.parse_args()
    levels = list(orc.LEVELS) if a.levels == 'all' else a.levels.split(',')
    for lv in levels:
        ep = 0
        for s in range(a.seeds):
            ep += compare(lv, 100 + s, a.steps, a.policy, act_seed=s)
        print('%20s OK  (%d seeds x %d steps, %d episodes)' % (lv, a.seeds, a.steps, ep), flush=True)
