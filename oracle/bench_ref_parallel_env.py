"""CPU baseline of the REFERENCE's own vectoriser: unmodified
babyai.rl.utils.penv.ParallelEnv (penv.py:18-59, one forked process per env)
over envs built exactly like scripts/train_rl.py:53-60, timed around
ParallelEnv.step only.  The MiniGrid core underneath is the clean-room shim
(oracle/shim) because the third-party package is absent, so the number is
representative of, not identical to, the original.  Build container only.

usage: python oracle/bench_ref_parallel_env.py [--level GoToRedBall] [--procs 64] [--steps 1000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refenv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--level', default='GoToRedBall')
    ap.add_argument('--procs', type=int, default=64)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=100)
    ap.add_argument('--rng', default='mt')
    a = ap.parse_args()
    gym = refenv.setup(a.rng)
    from babyai.rl.utils.penv import ParallelEnv
    envs = []
    for i in range(a.procs):
        env = gym.make('BabyAI-%s-v0' % a.level)
        env.seed(100 * 1 + i)
        envs.append(env)
    # single-process, no IPC
    e0 = gym.make('BabyAI-%s-v0' % a.level)
    e0.seed(100)
    e0.reset()
    acts = np.random.RandomState(0).randint(0, 7, (a.steps + a.warmup, a.procs))
    t0 = time.perf_counter()
    for t in range(a.steps):
        _, _, d, _ = e0.step(acts[t, 0])
        if d:
            e0.reset()
    single = a.steps / (time.perf_counter() - t0)
    penv = ParallelEnv(envs)
    penv.reset()
    for t in range(a.warmup):
        list(penv.step(acts[t]))
    t0 = time.perf_counter()
    for t in range(a.warmup, a.warmup + a.steps):
        list(penv.step(acts[t]))
    dt = time.perf_counter() - t0
    print(json.dumps({'level': a.level, 'procs': a.procs, 'steps': a.steps, 'host_cores': os.cpu_count(),
                      'parallel_env_steps_per_s': a.procs * a.steps / dt, 'single_process_steps_per_s': single,
                      'minigrid_core': 'clean-room shim (oracle/shim)', 'rng': a.rng}))


if __name__ == '__main__':
    main()
