"""Generates tests/golden/*.npz by running the REFERENCE's own BabyAI layer
(/root/reference/babyai/levels/*, unmodified) on the gym_minigrid shim
(oracle/shim) with the Philox back-end, under ParallelEnv's auto-reset rule
(penv.py:7-11).  Build container only (needs /root/reference); the committed
.npz files travel to the GPU box.

Per level: K traces of T steps.  Actions are 75 % reference-bot (babyai/bot.py)
/ 25 % uniform random so that episodes actually succeed and the pickup / drop /
toggle / PutNext / Before / After paths of the verifier are exercised.

usage: python tests/golden/make_golden.py [--only-missing] [--success]
       BABYAI_DONE_ACTIONS=1 python tests/golden/make_golden.py --done-actions
       python tests/golden/make_golden.py --bonus [--only-missing]
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, '..', '..'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import refenv  # noqa: E402

CONFIG_LEVELS = ['GoToRedBall', 'GoToLocal', 'PickupLoc', 'GoTo', 'BossLevel']
OTHER_LEVELS = ['GoToRedBallGrey', 'GoToObjMazeS4R2', 'GoToOpen', 'Pickup', 'GoToSeq', 'Synth', 'SynthSeq',
                'MiniBossLevel', 'BossLevelNoUnlock', 'Open', 'PutNext', 'PutNextLocal', 'PutNextLocalS5N3', 'UnblockPickup']


def trace(level, seed, T, act_seed, p_bot=0.75, p_done=0.0):
    env = refenv.make_env(level, seed, 'philox')
    from babyai.bot import Bot
    rng = np.random.RandomState(act_seed)
    obs = env.reset()
    obs0 = obs['image'].copy().reshape(-1)
    dir0 = obs['direction']
    missions = [obs['mission']]
    bot, last = Bot(env), None
    A = np.zeros(T, np.int8)
    O = np.zeros((T, 147), np.uint8)
    R = np.zeros(T, np.float32)
    D = np.zeros(T, np.uint8)
    Q = np.zeros(T, np.int8)
    for t in range(T):
        a = None
        if p_done and rng.rand() < p_done:
            a = 6                 # the `done` action: how an instruction reports in BABYAI_DONE_ACTIONS mode
        if a is None and bot is not None and rng.rand() < p_bot:
            try:
                a = int(bot.replan(last))
            except Exception:
                bot = None
        if a is None:
            a = int(rng.randint(0, 7))
            if bot is not None:   # things the bot cannot recover from (scripts/eval_bot.py:131-139)
                fc = env.grid.get(*env.front_pos)
                if a == 5 and fc is not None and (fc.type == 'box' or (fc.type == 'door' and fc.is_open)):
                    a = 6
        last = a
        obs, reward, done, _ = env.step(a)
        if done:
            obs = env.reset()
            missions.append(obs['mission'])
            bot, last = Bot(env), None
        A[t], R[t], D[t], Q[t] = a, np.float32(reward), done, obs['direction']
        O[t] = obs['image'].reshape(-1)
    return dict(actions=A, obs0=obs0, dir0=dir0, obs=O, reward=R, done=D, direction=Q, missions=missions)


# every other served level (added at the end of round 1): short traces, replayed by the CPU tests (oracle, host build of the
# kernel logic); tests/common.py lists which files the GPU suite replays
MORE_LEVELS = ['GoToRedBallNoDists', 'GoToObj', 'GoToObjS4', 'GoToObjS6', 'GoToLocalS5N2', 'GoToLocalS6N2', 'GoToLocalS6N3',
               'GoToLocalS6N4', 'GoToLocalS7N4', 'GoToLocalS7N5', 'GoToLocalS8N2', 'GoToLocalS8N3', 'GoToLocalS8N4',
               'GoToLocalS8N5', 'GoToLocalS8N6', 'GoToLocalS8N7', 'PutNextLocalS6N4', 'GoToObjMaze', 'GoToObjMazeOpen',
               'GoToObjMazeS4', 'GoToObjMazeS5', 'GoToObjMazeS6', 'GoToObjMazeS7', 'GoToSeqS5R2', 'SynthLoc', 'SynthS5R2',
               'GoToImpUnlock', 'Unlock']


# success-heavy traces (round 2): 97 % reference-bot actions, long enough for >= 50 SUCCESSFUL episodes
# per level, so that the PutNext / Before / After / And success paths of the verifier (verifier.py:393-550), the key /
# locked-door paths and the fp64 reward are pinned by reference-generated vectors that the CUDA pool replays
SUCCESS_LEVELS = {'PutNext': (8, 900), 'SynthSeq': (8, 1500), 'GoToSeq': (8, 1300), 'BossLevel': (8, 1800), 'MiniBossLevel': (8, 700),
                  'Unlock': (8, 850), 'GoToImpUnlock': (8, 1100), 'PutNextLocal': (4, 500), 'PickupLoc': (4, 300)}


def save(out, seeds, tr):
    np.savez_compressed(
        out, seeds=np.array(seeds, np.uint64),
        actions=np.stack([t['actions'] for t in tr]), obs0=np.stack([t['obs0'] for t in tr]),
        dir0=np.array([t['dir0'] for t in tr], np.int8), obs=np.stack([t['obs'] for t in tr]),
        reward=np.stack([t['reward'] for t in tr]), done=np.stack([t['done'] for t in tr]),
        direction=np.stack([t['direction'] for t in tr]),
        missions=np.array(json.dumps([t['missions'] for t in tr])))


def main_success():
    only_missing = '--only-missing' in sys.argv
    for level, (K, T) in SUCCESS_LEVELS.items():
        out = os.path.join(HERE, 'succ_' + level + '.npz')
        if only_missing and os.path.exists(out):
            continue
        seeds = [5000 + 31 * k for k in range(K)]
        tr = [trace(level, s, T, act_seed=100 + k, p_bot=0.97) for k, s in enumerate(seeds)]
        save(out, seeds, tr)
        eps = sum(int(t['done'].sum()) for t in tr)
        succ = sum(int((t['reward'] > 0).sum()) for t in tr)
        print('%20s  %d traces x %d steps, %d episodes (%d successes) -> %s (%d KB)'
              % (level, K, T, eps, succ, os.path.basename(out), os.path.getsize(out) // 1024), flush=True)


# verifier.use_done_actions (BABYAI_DONE_ACTIONS=1, read when babyai.levels.verifier is imported): the instructions only
# report through the `done` action -- 'success' if the previous action completed them, 'failure' otherwise
DONE_LEVELS = {'GoToLocal': (6, 500), 'PickupLoc': (6, 400), 'PutNextLocal': (6, 600), 'GoToSeqS5R2': (6, 700), 'SynthS5R2': (6, 700),
               'MiniBossLevel': (6, 800), 'Open': (4, 600)}


def main_done():
    assert os.environ.get('BABYAI_DONE_ACTIONS'), 'run with BABYAI_DONE_ACTIONS=1'
    for level, (K, T) in DONE_LEVELS.items():
        out = os.path.join(HERE, 'done_' + level + '.npz')
        seeds = [9000 + 13 * k for k in range(K)]
        tr = [trace(level, s, T, act_seed=300 + k, p_bot=0.8, p_done=0.10) for k, s in enumerate(seeds)]
        save(out, seeds, tr)
        eps = sum(int(t['done'].sum()) for t in tr)
        succ = sum(int((t['reward'] > 0).sum()) for t in tr)
        print('%20s  %d traces x %d steps, %d episodes (%d successes) -> %s (%d KB)'
              % (level, K, T, eps, succ, os.path.basename(out), os.path.getsize(out) // 1024), flush=True)


def main_bonus():
    """the 50 levels of babyai/levels/bonus_levels.py: 3 traces x 300 steps each, 60 % bot / 40 % random actions (random only
    where the reference's bot does not return)"""
    sys.path.insert(0, ROOT)
    from babyai_b200.levels import BONUS_LEVELS
    only_missing = '--only-missing' in sys.argv
    for level in BONUS_LEVELS:
        out = os.path.join(HERE, 'bonus_' + level + '.npz')
        if only_missing and os.path.exists(out):
            continue
        K, T = 3, 300
        seeds = [2000 + 11 * k for k in range(K)]
        p_bot = 0.0 if level in ('UnlockToUnlock', 'KeyInBox') else 0.6
        tr = [trace(level, s, T, act_seed=500 + k, p_bot=p_bot) for k, s in enumerate(seeds)]
        save(out, seeds, tr)
        eps = sum(int(t['done'].sum()) for t in tr)
        succ = sum(int((t['reward'] > 0).sum()) for t in tr)
        print('%24s  %d traces x %d steps, %d episodes (%d successes) -> %s (%d KB)'
              % (level, K, T, eps, succ, os.path.basename(out), os.path.getsize(out) // 1024), flush=True)


def main():
    if '--bonus' in sys.argv:
        return main_bonus()
    if '--success' in sys.argv:
        return main_success()
    if '--done-actions' in sys.argv:
        return main_done()
    only_missing = '--only-missing' in sys.argv
    for level in CONFIG_LEVELS + OTHER_LEVELS + MORE_LEVELS:
        if only_missing and os.path.exists(os.path.join(HERE, level + '.npz')):
            continue
        K, T = (4, 400) if level in CONFIG_LEVELS else (2, 300)
        if level == 'GoTo':
            T = 700
        if level == 'BossLevel':
            K, T = 6, 1500
        if level in ('SynthSeq', 'MiniBossLevel', 'BossLevelNoUnlock', 'Open', 'PutNext', 'UnblockPickup', 'GoToImpUnlock', 'Unlock'):
            K, T = 3, 800
        seeds = [1000 + 17 * k for k in range(K)]
        tr = [trace(level, s, T, act_seed=k) for k, s in enumerate(seeds)]
        out = os.path.join(HERE, level + '.npz')
        np.savez_compressed(
            out, seeds=np.array(seeds, np.uint64),
            actions=np.stack([t['actions'] for t in tr]), obs0=np.stack([t['obs0'] for t in tr]),
            dir0=np.array([t['dir0'] for t in tr], np.int8), obs=np.stack([t['obs'] for t in tr]),
            reward=np.stack([t['reward'] for t in tr]), done=np.stack([t['done'] for t in tr]),
            direction=np.stack([t['direction'] for t in tr]),
            missions=np.array(json.dumps([t['missions'] for t in tr])))
        eps = sum(int(t['done'].sum()) for t in tr)
        succ = sum(int((t['reward'] > 0).sum()) for t in tr)
        print('%20s  %d traces x %d steps, %d episodes (%d successes) -> %s (%d KB)'
              % (level, K, T, eps, succ, os.path.basename(out), os.path.getsize(out) // 1024), flush=True)


if __name__ == '__main__':
    main()
