"""The demo wire format (scripts/make_agent_demos.py:116, babyai/utils/demos.py:38-64): demos built from batched frozen-at-done
episodes must be what the reference's own one-env-at-a-time loop produces for the same seeds and actions, and must read
back through the reference's `transform_demos`."""
import os
import pickle
import sys

import numpy as np
import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))


def _batched_episodes(level, seeds, acts):
    """N oracle envs stepped together without auto-reset (ManyEnvs flavour), the records DemoRecorder collects"""
    import oracle as orc
    n = len(seeds)
    pool = orc.OraclePool(level, n, np.array(seeds, dtype=np.uint64))
    obs = pool.reset().copy()
    missions = [pool.mission(i) for i in range(n)]
    T = acts.shape[0]
    imgs, dirs, dones, rews = [], [], [], []
    for t in range(T):
        imgs.append(obs.copy()); dirs.append(np.asarray(pool.direction).copy())
        o, r, d = pool.step(acts[t], autoreset=False)
        obs = o.copy(); dones.append(d.copy()); rews.append(r.copy())
    return missions, np.stack(imgs), np.stack(dirs), np.stack(dones), np.stack(rews)


@pytest.mark.reference
@pytest.mark.parametrize('level', ['GoToLocal', 'PickupLoc', 'GoToObjMazeS4R2'])
def test_demos_equal_the_reference_loop(level):
    import refenv
    gym = refenv.setup('philox')
    import blosc                                              # the shim's stand-in (oracle/shim/blosc.py)
    from babyai.utils.demos import transform_demos
    from babyai_b200.demos import episodes_to_demos
    n, T, seed = 40, 150, 7000
    rng = np.random.RandomState(1)
    acts = rng.choice(7, size=(T, n), p=[0.15, 0.15, 0.45, 0.08, 0.05, 0.1, 0.02]).astype(np.int8)
    seeds = [seed + k for k in range(n)]
    missions, imgs, dirs, dones, rews = _batched_episodes(level, seeds, acts)
    ours = episodes_to_demos(missions, imgs, dirs, acts, dones, rews, pack_array=blosc.pack_array)
    # the reference's loop (make_agent_demos.generate_demos), one env at a time, same seeds, same action streams
    theirs = []
    for k in range(n):
        env = refenv.make_env(level, seeds[k], 'philox')
        obs = env.reset()
        mission, images, directions, actions = obs['mission'], [], [], []
        done, reward, t = False, 0, 0
        while not done and t < T:
            a = int(acts[t, k])
            new_obs, reward, done, _ = env.step(a)
            actions.append(a); images.append(obs['image']); directions.append(obs['direction'])
            obs = new_obs
            t += 1
        if done and reward > 0:
            theirs.append((mission, blosc.pack_array(np.array(images)), directions, actions))
    assert len(ours) == len(theirs) > 0
    a, b = transform_demos(pickle.loads(pickle.dumps(ours))), transform_demos(theirs)
    for da, db in zip(a, b):
        assert len(da) == len(db)
        for (oa, aa, na), (ob, ab, nb) in zip(da, db):
            assert np.array_equal(oa['image'], ob['image']) and oa['direction'] == ob['direction'] and oa['mission'] == ob['mission']
            assert aa == ab and na == nb


def test_unfinished_and_failed_episodes_are_dropped():
    from babyai_b200.demos import episodes_to_demos
    T, n = 5, 3
    imgs = np.zeros((T, n, 7, 7, 3), np.uint8)
    done = np.zeros((T, n), bool); rew = np.zeros((T, n), np.float32)
    done[2, 0] = True; rew[2, 0] = 0.5                      # success after 3 steps
    done[4, 1] = True                                       # timed out: reward 0
    demos = episodes_to_demos(['a', 'b', 'c'], imgs, np.zeros((T, n), np.int8), np.ones((T, n), np.int8), done, rew, pack_array=pickle.dumps)
    assert len(demos) == 1 and demos[0][0] == 'a' and demos[0][3] == [1, 1, 1] and pickle.loads(demos[0][1]).shape == (3, 7, 7, 3)
    assert len(episodes_to_demos(['a', 'b', 'c'], imgs, np.zeros((T, n), np.int8), np.ones((T, n), np.int8), done, rew,
                                 pack_array=pickle.dumps, successful_only=False)) == 2


@pytest.mark.gpu
def test_demo_recorder_on_the_gpu():
    """DemoRecorder (pool in ManyEnvs mode, batched policy) == the same episodes through the oracle"""
    import torch
    from babyai_b200.demos import DemoRecorder
    n, seed = 64, 4200
    rec = DemoRecorder('GoToLocal', n, pack_array=pickle.dumps)
    g = torch.Generator(device='cuda').manual_seed(0)
    stream = []

    def policy(image, direction, tokens):
        a = torch.randint(0, 7, (n,), device='cuda', generator=g, dtype=torch.int8)
        a = torch.where(a >= 3, torch.full_like(a, 2), a)          # left / right / forward: GoTo missions succeed by wandering
        stream.append(a.cpu().numpy())
        return a
    demos = rec.record(policy, [seed + k for k in range(n)], max_steps=70)
    acts = np.stack(stream)
    missions, imgs, dirs, dones, rews = _batched_episodes('GoToLocal', [seed + k for k in range(n)], acts)
    from babyai_b200.demos import episodes_to_demos
    want = episodes_to_demos(missions, imgs, dirs, acts, dones, rews, pack_array=pickle.dumps)
    assert len(demos) == len(want) > 5
    for a, b in zip(demos, want):
        assert a[0] == b[0] and a[2] == b[2] and a[3] == b[3] and np.array_equal(pickle.loads(a[1]), pickle.loads(b[1]))
