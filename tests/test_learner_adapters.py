"""Device-resident consumer adapters (babyai_b200/learner.py) in the GPU-less container: the
facade runs over the host build of the kernel logic (tests/hostemu) through the `pool=` test hook, with CPU tensors
standing in for HBM.

test_reference_ppo_consumes_the_pool_unchanged is the "babyai/rl consumes it unchanged" check of the north star: the
reference's UNMODIFIED PPOAlgo / BaseAlgo / ACModel (imported from /root/reference, build container only) are run
twice from identical weights and torch seeds -- once the stock way (reference ParallelEnv forking reference envs on
the gym_minigrid shim + reference ObssPreprocessor), once on DeviceParallelEnv + learner.ObssPreprocessor after the
one-line rebinding INTEGRATION.md documents -- and must produce bit-identical experiences and parameter updates."""
import numpy as np
import pytest
import torch

import hostemu
import oracle as orc


@pytest.fixture
def single_torch_thread():
    """The reference's ParallelEnv forks its workers from this (multi-threaded) process; with OpenMP worker threads alive
    the parent's next backward pass was seen to stall.  One intra-op thread keeps the comparison deterministic and quick."""
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)

from babyai_b200 import ParallelEnv, make_envs
from babyai_b200.learner import DeviceManyEnvs, DeviceParallelEnv, FixedVocabulary, ObsBatch, ObssPreprocessor
from babyai_b200.levels import VOCAB, detokenize, level_spec


class EmuTensorPool(object):
    """BabyAIVecEnv's tensor interface over tests/hostemu (CPU tensors)."""

    def __init__(self, level, seeds, mode=0):
        n = len(seeds)
        self.emu = hostemu.HostEmuPool(level_spec(level), n, np.asarray(seeds, dtype=np.uint64), mode)
        self.num_envs = n
        self.device = torch.device('cpu')
        self.mission_tokens = torch.zeros((n, 72), dtype=torch.int16)
        self.direction = torch.zeros(n, dtype=torch.int8)
        self.reward = torch.zeros(n, dtype=torch.float32)
        self.done = torch.zeros(n, dtype=torch.uint8)

    def _tokens(self, idx):
        for i in idx:
            self.mission_tokens[i] = torch.from_numpy(self.emu.tokens(int(i)))

    def reset(self, obs=None, direction=None):
        obs.copy_(torch.from_numpy(self.emu.reset()))
        self.direction.copy_(torch.from_numpy(self.emu.direction))
        self._tokens(range(self.num_envs))
        return obs

    def step(self, actions, obs=None, reward=None, done=None, direction=None):
        assert actions.dtype == torch.int8 and actions.shape == (self.num_envs,)
        o, r, d = self.emu.step(actions.numpy())
        obs.copy_(torch.from_numpy(o))
        self.reward.copy_(torch.from_numpy(r))
        self.done.copy_(torch.from_numpy(d))
        self.direction.copy_(torch.from_numpy(self.emu.direction))
        self._tokens(np.nonzero(d)[0])
        return obs, self.reward, self.done


    def step_learner(self, actions, obs, reward, done, direction=None):
        """bb_pool_step_learner: host actions -> device obs / direction, host reward / done"""
        assert actions.dtype == np.int8 and reward.dtype == np.float32 and done.dtype == np.uint8
        o, r, d = self.emu.step(actions)
        obs.copy_(torch.from_numpy(o))
        reward[...], done[...] = r, d
        self.direction.copy_(torch.from_numpy(self.emu.direction))
        if direction is not None:
            direction.copy_(self.direction)
        self._tokens(np.nonzero(d)[0])

    # host-buffer interface (bb_pool_seed / reset_host / step_host / missions)
    def seed(self, seeds):
        self.emu.seed(np.asarray(list(seeds), dtype=np.uint64))

    def reset_host(self, obs, direction):
        obs[...] = self.emu.reset()
        direction[...] = self.emu.direction

    def step_host(self, actions, obs, reward, done, direction):
        o, r, d = self.emu.step(np.asarray(actions, dtype=np.int8))
        obs[...], reward[...], done[...], direction[...] = o, r, d, self.emu.direction

    def missions(self, idx=None):
        return [detokenize(self.emu.tokens(int(i))) for i in (range(self.num_envs) if idx is None else idx)]


def _device_env(level, n, seed=1, fused_io=False):
    envs = make_envs(level, n, seed=seed)
    return DeviceParallelEnv(envs, pool=EmuTensorPool(level, envs.seeds), fused_io=fused_io)


def _ref_tokens(mission, width):
    """InstructionsPreprocessor.__call__ (utils/format.py:59-75) with the fixed vocabulary."""
    import re
    ids = [VOCAB.index(w) for w in re.findall('([a-z]+)', mission.lower())]
    return ids + [0] * (width - len(ids))


@pytest.mark.parametrize('level,fused_io', [('GoToLocal', False), ('PutNextLocal', True), ('GoToSeqS5R2', False), ('GoToLocal', True)])
def test_device_env_and_preprocessor_against_oracle(level, fused_io):
    n, T = 12, 70
    env = _device_env(level, n, fused_io=fused_io)
    assert env.fused_io == fused_io
    o = orc.OraclePool(level, n, np.array([100 + i for i in range(n)], dtype=np.uint64))
    pre = ObssPreprocessor()
    obs = env.reset()
    assert isinstance(obs, ObsBatch) and len(obs) == n
    assert np.array_equal(obs.image.numpy(), o.reset())
    rng = np.random.RandomState(2)
    history, want_img, want_mis = [], [], []
    for t in range(T):
        # what BaseAlgo does with an observation batch before acting (base.py:134)
        p = pre(obs, device='cpu')
        assert p.image.dtype == torch.float32 and p.image.shape == (n, 7, 7, 3) and p.instr.dtype == torch.long
        missions = [o.mission(i) for i in range(n)]
        width = max(len(m.replace(',', '').split()) for m in missions)
        assert p.instr.shape == (n, width)
        assert p.instr.tolist() == [_ref_tokens(m, width) for m in missions]
        assert [obs[i]['mission'] for i in range(n)] == missions          # host view of single observations
        assert all(obs[i]['direction'] == int(o.direction[i]) for i in range(n))
        history.append(obs)
        want_img.append(p.image.clone())
        want_mis.append(missions)
        act = rng.randint(0, 7, n)
        obs, rew, done, info = env.step(act if t % 2 else torch.as_tensor(act))     # numpy and tensor actions
        oo, orr, od = o.step(act.astype(np.int8))
        assert np.array_equal(obs.image.numpy(), oo)
        assert np.array_equal(np.asarray(rew, dtype=np.float32).view(np.uint32), orr.view(np.uint32))
        assert np.array_equal(np.asarray(done), od.astype(bool)) and len(info) == n and info[0] == {}
        assert torch.tensor(rew).dtype == torch.float32                    # base.py:158,167,175
    # the env-major list BaseAlgo hands the preprocessor after a rollout (base.py:208-210)
    flat = [history[i][j] for j in range(n) for i in range(T)]
    p = pre(flat, device='cpu')
    width = max(len(m.replace(',', '').split()) for ms in want_mis for m in ms)
    assert p.image.shape == (n * T, 7, 7, 3) and p.instr.shape == (n * T, width)
    k = 0
    for j in range(n):
        for i in range(T):
            assert torch.equal(p.image[k], want_img[i][j])
            assert p.instr[k].tolist() == _ref_tokens(want_mis[i][j], width)
            k += 1
    # row indexing as PPOAlgo does it (ppo.py: exps[inds + i] -> sb.obs)
    sb = p[np.array([3, 5, 8])]
    assert sb.image.shape == (3, 7, 7, 3) and torch.equal(sb.image[1], p.image[5]) and len(sb) == 3
    with pytest.raises(NotImplementedError):
        env.render()


def test_fixed_vocabulary(tmp_path):
    import json
    v = FixedVocabulary(str(tmp_path / 'model' / 'vocab.json'))
    assert v['go'] == 1 and v.max_size == 100 and len(v.vocab) == 32
    v.save()
    assert json.load(open(str(tmp_path / 'model' / 'vocab.json'))) == v.vocab
    assert all(detokenize([v[w] for w in m.replace(',', '').split()]) == m
               for m in ['go to the red ball', 'pick up a key, then open the door'])
    with pytest.raises(KeyError):
        v['lava']


@pytest.mark.reference
@pytest.mark.timeout(300)
@pytest.mark.parametrize('level', ['PutNextLocal', 'GoToObjMazeS4R2'])
def test_reference_ppo_consumes_the_pool_unchanged(level, single_torch_thread):
    import refenv
    gym = refenv.setup('philox')
    import babyai.rl
    import babyai.rl.algos.base as base
    import babyai.utils as utils
    from babyai.model import ACModel
    n, T, seed = 6, 40, 1
    # arch without 'res': the reference's residual FiLM block does `out += x` on a ReLU output (model.py:249), which
    # autograd of torch >= 1.5 rejects -- a modern-torch issue of the consumer, independent of the environment side
    # --- stock pipeline: scripts/train_rl.py:53-60, 86-114 ------------------------------------------------------
    envs = []
    for i in range(n):
        env = gym.make('BabyAI-%s-v0' % level)
        env.seed(100 * seed + i)
        envs.append(env)
    ref_pre = utils.ObssPreprocessor('b200-test-model', envs[0].observation_space)
    ref_pre.vocab.vocab = dict(FixedVocabulary().vocab)           # what loading the saved vocab.json gives (format.py:19-20)
    torch.manual_seed(0)
    model_a = ACModel(ref_pre.obs_space, envs[0].action_space, 128, 128, 128, True, 'gru', True, 'bow_endpool')
    # --- the same with the pool ---------------------------------------------------------------------------------
    pool_envs = make_envs(level, n, seed=seed)
    our_pre = ObssPreprocessor(dictlist=babyai.rl.DictList)
    assert our_pre.obs_space == ref_pre.obs_space
    torch.manual_seed(0)
    model_b = ACModel(our_pre.obs_space, pool_envs[0].action_space, 128, 128, 128, True, 'gru', True, 'bow_endpool')
    model_b.load_state_dict(model_a.state_dict())
    reshape = lambda _0, _1, reward, _2: 20 * reward              # noqa: E731  train_rl.py:110 with the default reward_scale

    def run(envs, model, pre, parallel_env):
        stock = base.ParallelEnv
        base.ParallelEnv = parallel_env                            # the one rebinding
        try:
            torch.manual_seed(7)
            np.random.seed(7)                                      # ppo.py shuffles the batch starts with numpy's global stream
            algo = babyai.rl.PPOAlgo(envs, model, T, 0.99, 1e-4, 0.9, 0.999, 0.99, 0.01, 0.5, 0.5, 20, 1e-5, 0.2, 2, 40, pre, reshape)
        finally:
            base.ParallelEnv = stock
        out = []
        for _ in range(2):
            exps, logs = algo.collect_experiences()
            out.append((exps, logs))
        logs = algo.update_parameters()
        return out, logs

    out_a, logs_a = run(envs, model_a, ref_pre, base.ParallelEnv)
    # the host-buffer facade (obs dicts with mission strings) under the reference's own preprocessor
    model_c = ACModel(ref_pre.obs_space, envs[0].action_space, 128, 128, 128, True, 'gru', True, 'bow_endpool')
    model_c.load_state_dict(model_b.state_dict())
    out_c, logs_c = run(pool_envs, model_c, ref_pre, lambda e: ParallelEnv(e, pool=EmuTensorPool(level, e.seeds)))
    out_b, logs_b = run(pool_envs, model_b, our_pre,
                        lambda e: DeviceParallelEnv(e, pool=EmuTensorPool(level, e.seeds)))
    for out_x, logs_x, model_x in ((out_b, logs_b, model_b), (out_c, logs_c, model_c)):
        for (ea, la), (eb, lb) in zip(out_a, out_x):
            assert torch.equal(ea.obs.image, eb.obs.image) and torch.equal(ea.obs.instr, eb.obs.instr)
            for k in ('action', 'reward', 'value', 'advantage', 'returnn', 'log_prob', 'mask', 'memory'):
                assert torch.equal(getattr(ea, k), getattr(eb, k)), k
            assert la['return_per_episode'] == lb['return_per_episode'] and la['num_frames'] == lb['num_frames']
        assert logs_a['policy_loss'] == logs_x['policy_loss'] and logs_a['grad_norm'] == logs_x['grad_norm']
        for pa, pb in zip(model_a.parameters(), model_x.parameters()):
            assert torch.equal(pa, pb)


@pytest.mark.reference
@pytest.mark.timeout(300)
@pytest.mark.parametrize('level', ['GoToLocal', 'PickupLoc'])
def test_reference_batch_evaluate_consumes_the_pool_unchanged(level, single_torch_thread):
    """babyai.evaluate.batch_evaluate (evaluate.py:85-140), unmodified, with `ManyEnvs` rebound."""
    import refenv
    refenv.setup('philox')
    import babyai.evaluate as evaluate
    import babyai.utils as utils
    from babyai.model import ACModel
    from babyai_b200 import ManyEnvs
    pre = utils.ObssPreprocessor('b200-test-model')
    pre.vocab.vocab = dict(FixedVocabulary().vocab)
    torch.manual_seed(3)
    model = ACModel(pre.obs_space, make_envs(level, 1)[0].action_space, 128, 128, 128, True, 'gru', True, 'bow_endpool')
    model.eval()

    def run(many_envs, preproc=None):
        stock = evaluate.ManyEnvs
        evaluate.ManyEnvs = many_envs
        try:
            torch.manual_seed(11)
            agent = utils.ModelAgent(model, preproc or pre, argmax=False)
            return evaluate.batch_evaluate(agent, 'BabyAI-%s-v0' % level, 10 ** 9, 20)    # 2 chunks of 10 envs, val seeds
        finally:
            evaluate.ManyEnvs = stock

    a = run(evaluate.ManyEnvs)
    b = run(lambda envs: ManyEnvs(envs, pool=EmuTensorPool(level, [0] * len(envs), mode=1)))
    # observations resident on the "device": DeviceManyEnvs + learner.ObssPreprocessor as the agent's preprocessor
    import babyai.rl
    c = run(lambda envs: DeviceManyEnvs(envs, pool=EmuTensorPool(level, [0] * len(envs), mode=1)),
            ObssPreprocessor(dictlist=babyai.rl.DictList))
    assert [int(x) for x in a['num_frames_per_episode']] == [int(x) for x in c['num_frames_per_episode']]
    assert [np.float32(x) for x in a['return_per_episode']] == [np.float32(x) for x in c['return_per_episode']]
    assert list(a['seed_per_episode']) == list(b['seed_per_episode'])
    assert [int(x) for x in a['num_frames_per_episode']] == [int(x) for x in b['num_frames_per_episode']]
    # the pool hands out the reward as float32 -- the precision the learner consumes it in (base.py:167,175); the
    # reference env returns the Python double before that rounding
    assert [np.float32(x) for x in a['return_per_episode']] == [np.float32(x) for x in b['return_per_episode']]
    assert max(a['return_per_episode']) > 0 or level != 'GoToLocal'


@pytest.mark.parametrize('level', ['GoToRedBall', 'PickupLoc'])
def test_single_env_gym_surface(level):
    """babyai_b200.gymapi: gym.make / seed / reset / step / spaces / actions of ONE environment (imitation.py:84,114,
    scripts/enjoy.py:35-44), registered under the reference's ids with the shim's gym."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(orc.__file__)), 'shim'))
    import gym
    from babyai_b200 import gymapi
    env = gymapi.make('BabyAI-%s-v0' % level, pool=EmuTensorPool(level, [0], mode=1))
    assert env.action_space.n == 7 and env.observation_space.spaces['image'].shape == (7, 7, 3)
    assert env.actions.forward == 2 and env.actions.done == 6 and env.level_name == level and env.unwrapped is env
    o = orc.OraclePool(level, 1, np.array([4242], dtype=np.uint64))
    assert env.seed(4242) == [4242]
    rng = np.random.RandomState(5)
    for episode in range(4):                              # reset() continues the random stream: next level of the seed
        obs = env.reset()
        assert np.array_equal(obs['image'], o.reset()[0]) and obs['mission'] == o.mission(0) == env.mission
        assert obs['direction'] == int(o.direction[0])
        for t in range(80):
            a = int(rng.randint(0, 7))
            obs, reward, done, info = env.step(a)
            oo, orr, od = o.step(np.array([a], dtype=np.int8), autoreset=False)
            assert np.array_equal(obs['image'], oo[0]) and np.float32(reward) == orr[0] and done == bool(od[0]) and info == {}
            if done:
                again = env.step(2)                       # a finished env repeats its terminal result until reset()
                assert np.array_equal(again[0]['image'], obs['image']) and again[1] == reward and again[2] is True
                break
        assert done                                       # max_steps = 64
    with pytest.raises(NotImplementedError):
        env.render()
    # registration under the reference's ids (levelgen.py:481-486)
    saved = dict(gym.envs.registration.registry.env_specs)          # other tests make the REFERENCE's envs by these ids
    try:
        ids = gymapi.register_levels(gym)
        assert 'BabyAI-BossLevel-v0' in ids and 'BabyAI-Unlock-v0' in ids and 'BabyAI-KeyCorridorS3R3-v0' in ids and len(ids) == 97      # the 47 ICLR-19 levels + the 50 bonus levels
        assert gym.spec('BabyAI-%s-v0' % level).entry_point.func is gymapi.SingleEnv
    finally:
        gym.envs.registration.registry.env_specs.clear()
        gym.envs.registration.registry.env_specs.update(saved)
