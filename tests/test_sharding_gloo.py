"""Host-side logic of the multi-GPU path, world_size 2 over gloo on CPU: the env
index space is split contiguously, seeds follow the global index, the counter
all-gather sums to the single-process totals, and each rank's slice reproduces
the matching slice of a single-process run."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))


def _worker(rank, world, port, total, steps, q):
    for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests', 'hostemu')):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import hostemu
    from babyai_b200.levels import level_spec
    from babyai_b200.sharding import gather_counters, shard_range, shard_seeds
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    seeds = shard_seeds(1, total, rank, world)
    pool = hostemu.HostEmuPool(level_spec('GoToLocal'), hi - lo, seeds)
    pool.reset()
    acts = np.random.RandomState(0).randint(0, 7, (steps, total)).astype(np.int8)
    eps, obs_sum = 0, 0
    for t in range(steps):
        o, r, d = pool.step(acts[t, lo:hi])
        eps += int(d.sum())
        obs_sum += int(o.astype(np.int64).sum())
    tot = gather_counters({'steps': (hi - lo) * steps, 'episodes': eps, 'successes': 0, 'errors': 0})
    q.put((rank, lo, hi, eps, obs_sum, tot))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_process():
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'hostemu'))
    import hostemu
    from babyai_b200.levels import level_spec
    from babyai_b200.sharding import shard_range, shard_seeds
    total, steps, world = 50, 80, 2
    assert [shard_range(7, r, 3) for r in range(3)] == [(0, 3), (3, 5), (5, 7)]
    # single process, all envs
    pool = hostemu.HostEmuPool(level_spec('GoToLocal'), total, shard_seeds(1, total, 0, 1))
    pool.reset()
    acts = np.random.RandomState(0).randint(0, 7, (steps, total)).astype(np.int8)
    eps_by_env = np.zeros(total, np.int64)
    obs_by_env = np.zeros(total, np.int64)
    for t in range(steps):
        o, r, d = pool.step(acts[t])
        eps_by_env += d
        obs_by_env += o.reshape(total, -1).astype(np.int64).sum(1)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, steps, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    for rank, lo, hi, eps, obs_sum, tot in res:
        assert eps == int(eps_by_env[lo:hi].sum())
        assert obs_sum == int(obs_by_env[lo:hi].sum())
        assert tot['steps'] == total * steps and tot['episodes'] == int(eps_by_env.sum())
