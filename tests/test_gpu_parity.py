"""Parity of the CUDA pool (through the C ABI) with the oracle and the golden
traces.  Bit-exact: observation bytes, float32 reward bit patterns, done,
direction, mission tokens, full hidden state."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from common import CONFIG_LEVELS, GOLDEN_LEVELS_GPU, compare_pools, replay_golden  # noqa: E402


class GpuPool(object):
    """Adapter: numpy in / numpy out around BabyAIVecEnv (device tensors)."""

    def __init__(self, level, n, seeds, mode=0, int64_actions=False):
        import torch
        from babyai_b200 import BabyAIVecEnv
        self.torch = torch
        self.env = BabyAIVecEnv(level, n, seeds=np.asarray(seeds, dtype=np.uint64), mode=mode)
        self.n = n
        self.int64 = int64_actions

    @property
    def direction(self):
        return self.env.direction.cpu().numpy()

    def reset(self):
        return self.env.reset().cpu().numpy()

    def step(self, actions):
        t = self.torch
        a = t.as_tensor(np.asarray(actions), device=self.env.device).to(t.int64 if self.int64 else t.int8).contiguous()
        o, r, d = self.env.step(a)
        return o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()

    def state(self, i):
        return self.env.state(i)

    def mission(self, i):
        return self.env.missions([i])[0]


@pytest.mark.parametrize('level', GOLDEN_LEVELS_GPU)
def test_gpu_replays_golden(level):
    replay_golden(level, lambda lv, n, s: GpuPool(lv, n, s), lambda p, i: p.mission(i))


@pytest.mark.parametrize('level,n,steps', [
    ('GoToRedBall', 64, 1000),      # BASELINE config 1: 64 envs, 1000 random-action steps
    ('GoToLocal', 2048, 300),
    ('PickupLoc', 2048, 300),
    ('GoTo', 512, 300),
    ('BossLevel', 512, 300),
    ('MiniBossLevel', 256, 400),
    ('SynthSeq', 128, 200),
    ('GoToObjMazeS4R2', 100, 300),  # ragged: not a multiple of 32
    ('PutNextLocal', 1024, 400),
    ('PutNextLocalS5N3', 512, 300),
    ('Open', 256, 300),
    ('PutNext', 256, 300),
    ('UnblockPickup', 128, 200),
])
def test_gpu_matches_oracle(level, n, steps):
    import oracle as orc
    # scripts/train_rl.py:59 seeding convention with --seed 1
    seeds = np.array([100 * 1 + i for i in range(n)], dtype=np.uint64)
    o = orc.OraclePool(level, n, seeds)
    g = GpuPool(level, n, seeds)
    # config 1 uses the shared action stream RandomState(0).randint(0, 7, (T, N))
    eps = compare_pools(o, g, n, steps, act_seed=0, state=(n <= 512),
                        mission_a=lambda p, i: p.mission(i), mission_b=lambda p, i: p.mission(i))
    assert eps > 0 or level in ('PutNext', 'UnblockPickup', 'Open')      # max_steps 576..1152: no episode may end in 300 steps
    assert g.env.counters()['errors'] == 0


@pytest.mark.parametrize('level,n', [('PickupLoc', 1024), ('PutNextLocal', 1024), ('MiniBossLevel', 256), ('BossLevel', 256)])
def test_gpu_matches_oracle_interaction_heavy_actions(level, n):
    """Pickup / drop / toggle heavy action mix (stale and refreshed obj_poss snapshots, opened boxes, toggled doors)."""
    import oracle as orc
    seeds = np.arange(n, dtype=np.uint64) * 3 + 77
    p = [0.12, 0.12, 0.30, 0.17, 0.14, 0.13, 0.02]
    g = GpuPool(level, n, seeds)
    compare_pools(orc.OraclePool(level, n, seeds), g, n, 400, act_seed=5, action_p=p, state=(n <= 256),
                  mission_a=lambda q, i: q.mission(i), mission_b=lambda q, i: q.mission(i))
    assert g.env.counters()['errors'] == 0


def test_preprocess_obss_adapter():
    """Device-resident stand-in for ObssPreprocessor (utils/format.py:100-119): float image + long token tensors."""
    import torch
    from babyai_b200 import BabyAIVecEnv, preprocess_obss, VOCAB
    env = BabyAIVecEnv('PickupLoc', 128, seeds=np.arange(128, dtype=np.uint64))
    env.reset()
    fn = preprocess_obss(env)
    out = fn(None, device='cuda')
    assert out.image.shape == (128, 7, 7, 3) and out.image.dtype == torch.float32 and out.image.is_cuda
    assert out.instr.dtype == torch.long and out.instr.shape[0] == 128 and int(out.instr.max()) < len(VOCAB)
    assert fn.vocab['pick'] == VOCAB.index('pick')
    from babyai_b200 import detokenize
    assert [detokenize(row.tolist()) for row in out.instr.cpu()] == env.missions()


def test_int64_actions_and_counters():
    import oracle as orc
    n = 96
    seeds = np.arange(n, dtype=np.uint64) + 5
    o = orc.OraclePool('GoToLocal', n, seeds)
    g = GpuPool('GoToLocal', n, seeds, int64_actions=True)
    eps = compare_pools(o, g, n, 100, act_seed=4, state=False)
    c = g.env.counters()
    assert c['steps'] == n * 100 and c['episodes'] == eps and c['errors'] == 0
    assert 0 < c['successes'] <= c['episodes']


def test_freeze_mode():
    """ManyEnvs flavour through the C ABI."""
    import oracle as orc
    level, n = 'PickupLoc', 64
    seeds = np.arange(n, dtype=np.uint64) + 900
    o = orc.OraclePool(level, n, seeds)
    g = GpuPool(level, n, seeds, mode=1)
    assert np.array_equal(o.reset(), g.reset())
    for i in range(n):
        assert o.state(i) [1] == g.state(i)[1]          # incl. draws / attempts: nothing pre-generated
    rng = np.random.RandomState(3)
    frozen = np.zeros(n, bool)
    last = [None] * n
    for t in range(70):
        act = rng.randint(0, 7, n).astype(np.int8)
        go, gr, gd = g.step(act)
        oo, orr, od = o.step(act, autoreset=False)
        for i in range(n):
            if frozen[i]:
                assert (go[i] == last[i][0]).all() and gr[i] == last[i][1] and gd[i]
            else:
                assert (go[i] == oo[i]).all() and gr[i] == orr[i] and gd[i] == od[i]
                if od[i]:
                    frozen[i] = True
                    last[i] = (oo[i].copy(), orr[i])
    assert frozen.all()      # max_steps = 64 < 70


def test_reseed_and_reset_like_batch_evaluate():
    """evaluate.py:104-108: env.seed(range(...)) then reset() for each chunk of episodes."""
    import oracle as orc
    n = 32
    g = GpuPool('GoToLocal', n, np.arange(n), mode=1)
    for chunk in range(3):
        seeds = np.arange(n, dtype=np.uint64) + 10 ** 9 + chunk * n
        g.env.seed(seeds)
        o = orc.OraclePool('GoToLocal', n, seeds)
        assert np.array_equal(o.reset(), g.reset())
        act = np.random.RandomState(chunk).randint(0, 7, n).astype(np.int8)
        oo, orr, od = o.step(act, autoreset=False)
        go, gr, gd = g.step(act)
        assert np.array_equal(oo, go) and np.array_equal(orr, gr) and np.array_equal(od, gd)


@pytest.mark.parametrize('level,n,T', [('GoToLocal', 4096, 24), ('GoToLocal', 1000, 40), ('PickupLoc', 200, 40), ('GoToObjS4', 256, 40),
                                        ('BossLevel', 512, 16), ('GoToObjMazeS4R2', 300, 40)])
def test_rollout_graph_equals_stepwise(level, n, T):
    """bb_pool_rollout (persistent kernel: fused generator warp with full and ragged CTAs, refill passes for 4x4
    rooms, 22x22 staging with k_gen beside it) == T x bb_pool_step."""
    import torch
    from babyai_b200 import BabyAIVecEnv
    seeds = np.arange(n, dtype=np.uint64) + 77
    a = BabyAIVecEnv(level, n, seeds=seeds)
    b = BabyAIVecEnv(level, n, seeds=seeds)
    acts = torch.randint(0, 7, (T, n), device='cuda', dtype=torch.int8)
    a.reset(); b.reset()
    obs = torch.zeros((T, n, 7, 7, 3), dtype=torch.uint8, device='cuda')
    rew = torch.zeros((T, n), device='cuda')
    done = torch.zeros((T, n), dtype=torch.uint8, device='cuda')
    for rep in range(3):            # later calls consume levels generated during the earlier ones
        a.rollout(acts, obs, rew, done)
        for t in range(T):
            o, r, d = b.step(acts[t])
            assert torch.equal(o, obs[t]) and torch.equal(r, rew[t]) and torch.equal(d, done[t]), (rep, t)
    assert a.counters()['errors'] == 0


@pytest.mark.parametrize('knobs', [
    {'BB_GEN_FUSED': '2', 'BB_GEN_BUDGET': '1'},         # fused generator warp, one round per launch: deficits carry over
    {'BB_GEN_FUSED': '2', 'BB_GEN_BUDGET': '1', 'BB_RING_DEPTH': '48'},       # ... with a shallow ring: the must-complete rule (< 2T levels left) kicks in
    {'BB_GEN_FUSED': '0', 'BB_GEN_BUDGET': '1'},         # refill passes (k_gen_scan + k_gen_small) between launches instead
    {'BB_GEN_FUSED': '0', 'BB_REFILL_EVERY': '1', 'BB_GEN_MIN_ACTIVE': '0'},
])
def test_rollout_with_suspended_generation(monkeypatch, knobs):
    """Bounded generation (a round budget per launch, sparse warps stop early) leaves rings with a deficit that later
    launches work off; rollouts must still equal the oracle bit for bit, whichever way the levels are supplied."""
    import torch
    import oracle as orc
    from babyai_b200 import BabyAIVecEnv
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    n, T, R = 512, 16, 12
    seeds = np.arange(n, dtype=np.uint64) + 4242
    env = BabyAIVecEnv('PickupLoc', n, seeds=seeds)
    o = orc.OraclePool('PickupLoc', n, seeds)
    assert np.array_equal(env.reset().cpu().numpy(), o.reset())
    obs = torch.zeros((T, n, 7, 7, 3), dtype=torch.uint8, device='cuda')
    rew = torch.zeros((T, n), device='cuda')
    done = torch.zeros((T, n), dtype=torch.uint8, device='cuda')
    rng = np.random.RandomState(11)
    for r in range(R):
        acts = rng.randint(0, 7, (T, n)).astype(np.int8)
        env.rollout(torch.as_tensor(acts, device='cuda'), obs, rew, done)
        ho, hr, hd = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        for t in range(T):
            oo, orr, od = o.step(acts[t])
            assert np.array_equal(ho[t], oo) and np.array_equal(hr[t].view(np.uint32), orr.view(np.uint32)) and np.array_equal(hd[t], od), (r, t)
    assert env.counters()['errors'] == 0
    # and the per-step entry point continues correctly from there
    a = rng.randint(0, 7, n).astype(np.int8)
    go, gr, gd = env.step(torch.as_tensor(a, device='cuda'))
    oo, orr, od = o.step(a)
    assert np.array_equal(go.cpu().numpy(), oo) and np.array_equal(gd.cpu().numpy(), od)


def test_size_independent_properties_at_full_size():
    """BASELINE config 2 size (65 536 envs): invariants that need no oracle run."""
    import torch
    from babyai_b200 import BabyAIVecEnv
    n = 65536
    env = BabyAIVecEnv('GoToLocal', n, seeds=np.array([100 + i for i in range(n)], dtype=np.uint64))
    obs = env.reset().clone()
    # determinism: a second pool with the same seeds produces identical bytes
    env2 = BabyAIVecEnv('GoToLocal', n, seeds=np.array([100 + i for i in range(n)], dtype=np.uint64))
    assert torch.equal(obs, env2.reset())
    # the agent's own view cell (3, 6) is 'empty' at reset; channel values stay in range
    assert (obs[:, 3, 6, 0] == 1).all() and (obs[..., 0] <= 7).all() and (obs[..., 1] <= 5).all() and (obs[..., 2] <= 2).all()
    gen = torch.Generator(device='cuda').manual_seed(0)
    tot_done = 0
    for t in range(70):
        a = torch.randint(0, 7, (n,), device='cuda', dtype=torch.int8, generator=gen)
        o, r, d = env.step(a)
        o2, r2, d2 = env2.step(a)
        assert torch.equal(o, o2) and torch.equal(r, r2) and torch.equal(d, d2)
        assert ((r > 0) <= (d > 0)).all() and (r <= 1).all() and (r >= 0).all()
        tot_done += int(d.sum())
    c = env.counters()
    assert c['steps'] == 70 * n and c['episodes'] == tot_done and c['errors'] == 0
    assert c['episodes'] >= n          # max_steps = 64 < 70: every env finished at least once


@pytest.mark.parametrize('zerocopy', ['1', '0', '2'])
def test_host_buffer_facades(monkeypatch, zerocopy):
    """ParallelEnv / ManyEnvs drop-ins: list-of-dict observations with mission strings.  bb_pool_step_host with
    reward / done / direction / actions over mapped page-locked memory (default), with copies only (0) and with the
    observations over mapped memory too (2)."""
    import oracle as orc
    monkeypatch.setenv('BB_HOST_ZEROCOPY', zerocopy)
    from babyai_b200 import ManyEnvs, ParallelEnv, make_envs
    n = 64
    envs = make_envs('GoToRedBall', n, seed=1)
    assert len(envs) == n and envs[0].action_space.n == 7 and envs[0].observation_space.spaces['image'].shape == (7, 7, 3)
    penv = ParallelEnv(envs)
    o = orc.OraclePool('GoToRedBall', n, np.array([100 + i for i in range(n)], dtype=np.uint64))
    obs = penv.reset()
    oo = o.reset()
    assert all((obs[i]['image'] == oo[i]).all() and obs[i]['mission'] == o.mission(i) for i in range(n))
    rng = np.random.RandomState(0)
    for t in range(100):
        act = rng.randint(0, 7, n)
        obs, rew, done, info = penv.step(act)
        oo, orr, od = o.step(act.astype(np.int8))
        for i in range(n):
            assert (obs[i]['image'] == oo[i]).all() and obs[i]['direction'] == o.direction[i]
            assert obs[i]['mission'] == o.mission(i) and np.float32(rew[i]) == orr[i] and done[i] == bool(od[i])
    me = ManyEnvs(make_envs('GoToRedBall', 8))
    me.seed(range(10 ** 9, 10 ** 9 + 8))
    o = orc.OraclePool('GoToRedBall', 8, np.arange(8, dtype=np.uint64) + 10 ** 9)
    obs = me.reset()
    assert all((obs[i]['image'] == o.reset()[i]).all() for i in range(1))
