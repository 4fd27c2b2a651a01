"""Multi-GPU: environments are independent, so a job of G ranks is G replicas of
the pool, each owning a contiguous slice of the global environment index; seeds
follow the GLOBAL index so results do not depend on G.  The only collective on
this path is the all-gather of the four int64 counters."""
import numpy as np


def shard_range(total_envs, rank, world):
    """Global env indices [lo, hi) owned by `rank` (contiguous, sizes differ by at most 1)."""
    base, rem = divmod(total_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_seeds(seed, total_envs, rank, world):
    """env i of the job is seeded 100*seed + i (scripts/train_rl.py:59), whichever rank owns it."""
    lo, hi = shard_range(total_envs, rank, world)
    return np.array([100 * seed + i for i in range(lo, hi)], dtype=np.uint64)


def gather_counters(counters, device=None, group=None):
    """Sum of the per-rank {steps, episodes, successes, errors} over the job: one all-gather of 4 int64."""
    import torch
    import torch.distributed as dist
    keys = ('steps', 'episodes', 'successes', 'errors')
    t = torch.tensor([int(counters[k]) for k in keys], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        parts = [torch.zeros_like(t) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, t, group=group)
        t = torch.stack(parts).sum(0)
    return dict(zip(keys, (int(x) for x in t.tolist())))
