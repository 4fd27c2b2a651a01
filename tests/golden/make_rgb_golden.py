"""tests/golden/rgb_obs.npz: (observation, picture) pairs from gym_minigrid.wrappers.RGBImgPartialObsWrapper (the oracle
shim's restatement) wrapped around the reference's own levels, as scripts/train_rl.py:54-58 does for the pixel
architectures.  Build container only (needs /root/reference).

usage: python tests/golden/make_rgb_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..', 'oracle'))
import refenv  # noqa: E402


def main():
    gym = refenv.setup('philox')
    from gym_minigrid.wrappers import RGBImgPartialObsWrapper
    from babyai.bot import Bot
    obs_l, rgb_l = [], []
    rng = np.random.RandomState(0)
    for level, steps in (('BossLevel', 260), ('Unlock', 120), ('PutNextLocal', 60), ('GoToRedBall', 40)):
        env = gym.make('BabyAI-%s-v0' % level)
        env.seed(77)
        w = RGBImgPartialObsWrapper(env)
        raw = env.reset()
        bot, last = Bot(env), None
        for t in range(steps):
            if t % 4 == 0:
                obs_l.append(raw['image'].copy())
                rgb_l.append(w.observation(raw)['image'].copy())
            a = None
            if bot is not None and rng.rand() < 0.8:
                try:
                    a = int(bot.replan(last))
                except Exception:
                    bot = None
            if a is None:
                a = int(rng.randint(0, 6))
                bot = None
            last = a
            raw, _r, done, _ = env.step(a)
            if done:
                raw = env.reset()
                bot, last = Bot(env), None
    obs, rgb = np.stack(obs_l), np.stack(rgb_l)
    out = os.path.join(HERE, 'rgb_obs.npz')
    np.savez_compressed(out, obs=obs, rgb=rgb)
    print('%d (obs, picture) pairs, %d distinct cell codes -> %s (%d KB)' % (len(obs), len(np.unique(obs.reshape(-1, 3), axis=0)),
                                                                         out, os.path.getsize(out) // 1024))


if __name__ == '__main__':
    main()
