"""Demonstrations in the reference's wire format out of pool rollouts.

The reference stores a demonstration as the tuple `(mission, blosc.pack_array(uint8[T, 7, 7, 3]), directions, actions)`
(scripts/make_agent_demos.py:116, scripts/make_human_demos.py), one per SUCCESSFUL episode; `images[i]` / `directions[i]`
are the observation the agent saw BEFORE `actions[i]`, and `babyai.utils.demos.transform_demos` (utils/demos.py:38-64)
turns the list back into (obs, action, done) triples for imitation learning (imitation.py:225-321).  make_agent_demos
generates them one env at a time: `env.seed(seed + len(demos))`, reset, act until done, keep if reward > 0.

Here a whole batch of episodes is generated at once on the GPU (ManyEnvs flavour: finished envs freeze):

  episodes_to_demos   pure numpy: [T, N] step records of N frozen-at-done episodes -> list of demo tuples
  DemoRecorder        seeds N envs, steps them with a batched policy until all are done, returns the demos of the
                      successful ones in seed order (what `generate_demos` returns for the same seeds when every
                      episode succeeds; failed seeds are skipped here, retried with a fresh level there)

`blosc` is a third-party dependency of the reference that this image does not have: pass `pack_array=` explicitly or have
`blosc` importable (the reference needs it anyway to read the file back)."""
import numpy as np

from .levels import detokenize


def _packer(pack_array):
    if pack_array is not None:
        return pack_array
    try:
        import blosc
    except ImportError as ex:
        raise ImportError('writing demos in the reference format needs `blosc.pack_array` (reference setup.py:13); '
                          'install blosc or pass pack_array=') from ex
    return blosc.pack_array


def episodes_to_demos(missions, images, directions, actions, done, reward, pack_array=None, successful_only=True):
    """missions: N strings (the mission at reset); images uint8 [T, N, 7, 7, 3] / directions [T, N]: what env k saw before
    its action at step t; actions [T, N]; done [T, N] / reward [T, N]: the step results (an env's first `done` ends its
    episode; later rows are ignored).  Returns [(mission, packed images [L, 7, 7, 3], [direction]*L, [action]*L)] for the
    episodes that ended (with reward > 0 unless successful_only=False), in env order."""
    pack = _packer(pack_array)
    images, directions, actions = np.asarray(images), np.asarray(directions), np.asarray(actions)
    done, reward = np.asarray(done).astype(bool), np.asarray(reward)
    T, N = actions.shape
    demos = []
    for k in range(N):
        ends = np.nonzero(done[:, k])[0]
        if len(ends) == 0:
            continue                                   # the episode did not finish within T steps
        L = int(ends[0]) + 1
        if successful_only and not reward[L - 1, k] > 0:
            continue
        demos.append((missions[k], pack(np.array(images[:L, k])), [int(d) for d in directions[:L, k]],
                      [int(a) for a in actions[:L, k]]))
    return demos


class DemoRecorder(object):
    """Batched `generate_demos` (scripts/make_agent_demos.py:70-135) on the pool."""

    def __init__(self, level, num_envs, device=0, pack_array=None):
        import torch
        from .vecenv import MODE_FREEZE, BabyAIVecEnv
        self.torch = torch
        self.env = BabyAIVecEnv(level, num_envs, device=device, mode=MODE_FREEZE)
        self.pack_array = pack_array

    def record(self, policy, seeds, max_steps=None, successful_only=True):
        """policy(image uint8[N,7,7,3], direction int8[N], tokens int16[N,L]) -> N actions (device tensors in, any
        integer tensor / array out).  seeds: N ints (make_agent_demos uses seed + k for the k-th demo)."""
        t = self.torch
        env = self.env
        n = env.num_envs
        env.seed(np.asarray(list(seeds), dtype=np.uint64))
        obs = env.reset()
        missions = env.missions()
        imgs, dirs, acts, dones, rews = [], [], [], [], []
        finished = t.zeros(n, dtype=t.bool, device=env.device)
        steps = 0
        while True:
            a = policy(obs, env.direction, env.mission_tokens)
            a = t.as_tensor(np.asarray(a) if not t.is_tensor(a) else a).to(device=env.device, dtype=t.int8).contiguous()
            imgs.append(obs.clone()); dirs.append(env.direction.clone()); acts.append(a.clone())
            obs, r, d = env.step(a)
            dones.append(d.clone()); rews.append(r.clone())
            finished |= d.bool()
            steps += 1
            if bool(finished.all()) or (max_steps is not None and steps >= max_steps):
                break
        stack = lambda xs: t.stack(xs).cpu().numpy()
        return episodes_to_demos(missions, stack(imgs), stack(dirs), stack(acts), stack(dones), stack(rews),
                                 pack_array=self.pack_array, successful_only=successful_only)
