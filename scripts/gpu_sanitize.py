"""compute-sanitizer driver: every kernel of the pool once, at CI size, in one process.

    compute-sanitizer --tool memcheck  python scripts/gpu_sanitize.py
    compute-sanitizer --tool racecheck python scripts/gpu_sanitize.py
    compute-sanitizer --tool synccheck python scripts/gpu_sanitize.py

Kernels launched: k_seed, k_gen_scan, k_gen_small, k_gen<false>, k_gen<true>, k_rollout<1> (fused with the generator warp,
and T = 1 as the per-step kernel), k_rollout<8>, k_rollout_cta<false>, k_rollout_cta<true>, k_step8<1>, k_step8<8>,
k_step8<1, true>, k_render_rgb.  Exits non-zero if a result differs between two identical pools (determinism) or an
error counter is set; the sanitizer's own findings are in its log."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, ROOT)
from babyai_b200 import BabyAIVecEnv  # noqa: E402


def drive(level, n, T, int64=False):
    seeds = np.arange(n, dtype=np.uint64) + 31
    outs = []
    for rep in range(2):
        env = BabyAIVecEnv(level, n, seeds=seeds)
        g = torch.Generator(device='cuda').manual_seed(3)
        acts = torch.randint(0, 7, (T, n), device='cuda', dtype=torch.int8, generator=g)
        obs = torch.zeros((T, n, 7, 7, 3), dtype=torch.uint8, device='cuda')
        rew, done = torch.zeros((T, n), device='cuda'), torch.zeros((T, n), dtype=torch.uint8, device='cuda')
        dirs = torch.zeros((T, n), dtype=torch.int8, device='cuda')
        first = env.reset().clone()
        for _ in range(3):
            env.rollout(acts, obs, rew, done, dirs)
        steps = []
        for t in range(6):
            a = acts[t].to(torch.int64) if int64 else acts[t]
            o, r, d = env.step(a)
            steps.append((o.clone(), r.clone(), d.clone()))
        pics = env.render_rgb(obs[:2])
        torch.cuda.synchronize()
        assert env.counters()['errors'] == 0, (level, env.counters())
        outs.append((first, obs.clone(), rew.clone(), done.clone(), dirs.clone(), pics, steps))
        env.close()
    a, b = outs
    same = all(bool(torch.equal(x, y)) for x, y in zip(a[:6], b[:6])) and all(
        bool(torch.equal(x, y)) for s, t in zip(a[6], b[6]) for x, y in zip(s, t))
    print('%-16s n=%d T=%d int64=%s deterministic=%s' % (level, n, T, int64, same), flush=True)
    return same


def main():
    ok = True
    ok &= drive('GoToLocal', 200, 16)                 # k_rollout fused + generator warp, T = 1 per-step kernel, k_gen_small
    ok &= drive('GoToObjS4', 96, 16, int64=True)      # refill passes (rooms smaller than 6x6), k_rollout<8>
    ok &= drive('BossLevel', 100, 12)                 # k_rollout_cta<false>, k_gen<false>, k_step8<1>
    ok &= drive('GoTo', 64, 12, int64=True)           # k_step8<8>
    ok &= drive('Unlock', 70, 12)                     # k_rollout_cta<true>, k_gen<true>, k_step8<1, true>
    print('sanitize driver:', 'OK' if ok else 'MISMATCH')
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
