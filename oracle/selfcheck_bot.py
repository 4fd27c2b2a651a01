"""Semantic self-check of the gym_minigrid shim: run the reference's GOFAI bot
(babyai/bot.py, unmodified) the way scripts/eval_bot.py:97-187 does.  The bot
plans from gen_obs_grid() visibility and asserts on inconsistencies, so a wrong
vis-mask geometry / door / pickup / drop rule / verifier coupling makes it
fail.  TEST INFRASTRUCTURE ONLY; needs /root/reference.

usage: python oracle/selfcheck_bot.py [--levels A,B] [--runs N] [--rng mt|philox]
"""
import argparse
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refenv  # noqa: E402


def run(levels, runs, rng, seed0=1):
    refenv.setup(rng)
    from babyai.levels import level_dict
    from babyai.bot import Bot
    ok = True
    for name in levels:
        succ = 0
        for r in range(runs):
            mission = level_dict[name](seed=seed0 + r)
            expert = Bot(mission)
            last = None
            steps = 0
            while True:
                action = expert.replan(last)
                obs, reward, done, info = mission.step(action)
                last = action
                steps += 1
                if done:
                    if reward > 0:
                        succ += 1
                    else:
                        assert steps == mission.max_steps
                    break
        print('%20s: %d/%d' % (name, succ, runs), flush=True)
        ok = ok and succ == runs
    return ok


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--levels', default='GoToRedBall,GoToLocal,PickupLoc,GoTo,BossLevel')
    ap.add_argument('--runs', type=int, default=50)
    ap.add_argument('--rng', default='mt')
    a = ap.parse_args()
    if a.levels == 'all':
        refenv.setup(a.rng)
        from babyai.levels import level_dict
        lv = list(level_dict.keys())
    else:
        lv = a.levels.split(',')
    sys.exit(0 if run(lv, a.runs, a.rng) else 1)
