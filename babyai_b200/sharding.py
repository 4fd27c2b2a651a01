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


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def pin_to_gpu_numa(local_rank):
    """Bind this process to the CPUs of the NUMA node its GPU hangs off, BEFORE the CUDA context and any page-locked
    buffer exist: the first touch of the pinned staging buffers (bb_pool_step_host's DMA targets) then lands in memory
    local to the GPU's PCIe root, and the thread that spins in cudaStreamSynchronize stays next to it.  One rank per
    GPU; on a two-socket HGX box GPUs 0-3 / 4-7 sit on different sockets.  Returns a small dict for the
    bench line, or None when the topology cannot be read (then nothing is changed)."""
    import os
    import subprocess
    try:
        sel = str(local_rank)
        cvd = os.environ.get('CUDA_VISIBLE_DEVICES')
        if cvd:
            ids = [x for x in cvd.split(',') if x]
            if local_rank < len(ids):
                sel = ids[local_rank]
        out = subprocess.run(['nvidia-smi', '--query-gpu=pci.bus_id', '--format=csv,noheader', '-i', sel],
                             capture_output=True, text=True, timeout=20).stdout.strip().splitlines()
        bus = out[0].strip().lower()
        if len(bus.split(':')[0]) == 8:                       # nvidia-smi prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read())
        if node < 0:
            return None
        cpus = _parse_cpulist(open('/sys/devices/system/node/node%d/cpulist' % node).read())
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {'numa_node': node, 'cpus': len(cpus), 'pci': bus}
    except Exception:
        return None
