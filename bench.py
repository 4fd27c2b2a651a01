#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched BabyAI pool (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE.json configs[1] -- BabyAI-GoToLocal-v0,
65 536 environments per GPU, uniform random actions, ParallelEnv auto-reset on.
A "step" is one environment step of all 65 536 environments of a rank.

Own arm, printed as ONE JSON line by rank 0:
  value     env-steps/sec, whole job, actions resident in HBM ([T, N] int8),
            observations written to a [T, N, 147] rollout buffer (385 MB > L2),
            K steps run as K/T calls of bb_pool_rollout (T = 40): the persistent
            stepping kernel k_rollout + the bounded level refill (k_gen_scan,
            k_gen_small); CUDA events, barrier + synchronize on both sides,
            max over ranks.
  e2e       the same metric through the reference-facing host-buffer call
            (bb_pool_step_host = what ParallelEnv.step returns to BaseAlgo):
            actions host->device and obs/reward/done/direction device->host
            inside the timed region, every step.
  roofline  dominant kernel k_rollout: 153 algorithmic bytes per env-step
            (147 obs + 4 reward + 1 done + 1 action) x N envs x T
            steps per launch / its CUDA-event duration (bb_pool_rollout_timed),
            against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the oracle's C port of the same path on this box's host cores.

--impl reference: the CPU arm (oracle C port, all host threads, same workload);
the reference itself is Python on a third-party package that is absent here
and /root/reference does not exist on the GPU box.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LEVEL = 'GoToLocal'
N_ENVS = 65536
CHUNK = int(os.environ.get("BENCH_CHUNK", "40"))   # rollout length per launch (= --frames-per-proc, arguments.py:42)
ALGO_BYTES_PER_STEP = 153       # 147 obs + 4 reward + 1 done + 1 action
FALLBACK_HBM_GBS = 6650.0       # /opt/skills/guides/B200_PROFILING.md fallback
METRIC = 'env-steps/sec at 65 536 envs (GoToLocal); obs bit-exact vs CPU ref'


def hbm_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured'
        except Exception:
            pass
    return FALLBACK_HBM_GBS, 'fallback'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[2 + k].lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                continue
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_port(steps_budget_s, n_envs, threads):
    """Oracle C port (oracle/babyai_oracle.c) on the host cores: GoToLocal, random actions, auto-reset."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import oracle as orc
    pool = orc.OraclePool(LEVEL, n_envs, np.array([100 + i for i in range(n_envs)], dtype=np.uint64))
    pool.reset()
    rng = np.random.RandomState(0)
    acts = rng.randint(0, 7, (64, n_envs)).astype(np.int8)
    for k in range(2):
        pool.step(acts[k], nthreads=threads)
    t0 = time.perf_counter()
    k = 0
    while True:
        pool.step(acts[k % 64], nthreads=threads)
        k += 1
        if time.perf_counter() - t0 > steps_budget_s:
            break
    dt = time.perf_counter() - t0
    return n_envs * k / dt, k


def run_reference(args, rank):
    """CPU arm: the oracle port with every host thread, same workload and metric."""
    if rank != 0:
        return
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import oracle as orc
    threads = os.cpu_count() or 1
    n = N_ENVS
    pool = orc.OraclePool(LEVEL, n, np.array([100 + i for i in range(n)], dtype=np.uint64))
    pool.reset()
    acts = np.random.RandomState(0).randint(0, 7, (64, n)).astype(np.int8)
    t0 = time.perf_counter()
    pool.step(acts[0], nthreads=threads)
    per_step = time.perf_counter() - t0
    # keep the whole run within ~2 minutes: if needed a step covers only the first `m` environments
    total = per_step * (args.steps + args.warmup)
    m = n
    if total > 120:
        m = max(1024, int(n * 120 / total) // 1024 * 1024)
        pool = orc.OraclePool(LEVEL, m, np.array([100 + i for i in range(m)], dtype=np.uint64))
        pool.reset()
        acts = acts[:, :m].copy()
    for k in range(args.warmup):
        pool.step(acts[k % 64], nthreads=threads)
    t0 = time.perf_counter()
    for k in range(args.steps):
        pool.step(acts[k % 64], nthreads=threads)
    dt = time.perf_counter() - t0
    v = m * args.steps / dt
    sample = '%d steps x %d envs (of %d) of %s, oracle C port, %d host threads' % (args.steps, m, n, LEVEL, threads)
    out = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'env-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps * (n / m),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': {'workload': 'BabyAI-%s-v0, %d envs, uniform random actions, auto-reset' % (LEVEL, n),
                   'envs_per_step_sampled': m},
        'cpu_baseline': {'value': v, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(out), flush=True)


def lanes2_probe(args):
    """Child process of the default bench run (own CUDA context: whatever happens here cannot touch the parent's numbers).
    k_rollout2 -- the experimental two-lanes-per-environment rollout kernel (BB_ROLLOUT_LANES=2, csrc/rollout2.cuh), written
    after round 1's GPU budget was spent -- against k_rollout in the same process: bit equality of three rollouts from the
    same seeds and actions, then the same timing loop for both.  Prints one JSON dict."""
    import numpy as np
    import torch
    from babyai_b200 import BabyAIVecEnv
    n, T = args.envs, CHUNK
    seeds = np.array([100 + i for i in range(n)], dtype=np.uint64)
    os.environ.pop('BB_ROLLOUT_LANES', None)
    pools = [BabyAIVecEnv(args.level, n, seeds=seeds)]
    os.environ['BB_ROLLOUT_LANES'] = '2'                 # read by bb_pool_create
    pools.append(BabyAIVecEnv(args.level, n, seeds=seeds))
    os.environ.pop('BB_ROLLOUT_LANES')
    dev = torch.device('cuda', 0)
    gen = torch.Generator(device=dev).manual_seed(5)
    acts = torch.randint(0, 7, (T, n), device=dev, dtype=torch.int8, generator=gen)
    bufs = [(torch.zeros((T, n, 7, 7, 3), dtype=torch.uint8, device=dev), torch.zeros((T, n), device=dev),
             torch.zeros((T, n), dtype=torch.uint8, device=dev), torch.zeros((T, n), dtype=torch.int8, device=dev)) for _ in pools]
    for p in pools:
        p.reset()
    equal = True
    for _ in range(3):
        for p, b in zip(pools, bufs):
            p.rollout(acts, *b)
        torch.cuda.synchronize()
        equal = equal and all(bool(torch.equal(x, y)) for x, y in zip(*bufs))
    out = {'bit_equal_to_k_rollout': equal, 'level': args.level, 'envs': n, 'steps_per_launch': T}
    for name, p, b in (('k_rollout', pools[0], bufs[0]), ('k_rollout2', pools[1], bufs[1])):
        for _ in range(5):
            p.rollout(acts, *b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 40
        for _ in range(reps):
            p.rollout(acts, *b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        kr = [p.rollout_timed(acts, *b)[0] for _ in range(8)][2:]
        out[name] = {'us_per_step': 1e3 * ms / T, 'env_steps_per_s': n * T / (ms * 1e-3), 'kernel_ms': sum(kr) / len(kr),
                     'errors': p.counters()['errors']}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=200)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--envs', type=int, default=N_ENVS, help='environments per GPU')
    ap.add_argument('--level', default=LEVEL)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--lanes2-probe', action='store_true', help='internal: the k_rollout2 child process')
    ap.add_argument('--no-probe', action='store_true', help='skip the k_rollout2 child process')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if args.lanes2_probe:
        lanes2_probe(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from babyai_b200 import BabyAIVecEnv

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the pool has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n, K, W = args.envs, args.steps, args.warmup
    # rank r owns global env indices [r*n, (r+1)*n); seeds follow the global index (train_rl.py:59, --seed 1)
    from babyai_b200.sharding import gather_counters, shard_seeds
    seeds = shard_seeds(1, world * n, rank, world)
    env = BabyAIVecEnv(args.level, n, seeds=seeds, device=local_rank)
    env.reset()

    T = CHUNK
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    actions = torch.randint(0, 7, (T, n), device=dev, dtype=torch.int8, generator=gen)
    obs = torch.empty((T, n, 7, 7, 3), dtype=torch.uint8, device=dev)       # 385 MB at n = 65 536: larger than L2
    rew = torch.empty((T, n), dtype=torch.float32, device=dev)
    done = torch.empty((T, n), dtype=torch.uint8, device=dev)
    dirs = torch.empty((T, n), dtype=torch.int8, device=dev)

    def run_steps(k):
        full, rem = divmod(k, T)
        for _ in range(full):
            env.rollout(actions, obs, rew, done, dirs)
        for t in range(rem):
            env.step(actions[t], obs[t], rew[t], done[t], dirs[t])

    # ---- device-resident throughput ------------------------------------------------
    run_steps(max(W, 3))
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = env.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run_steps(K)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = env.launches() - l0
    # keep the GPU busy a little longer so the clock sampler sees load even for short K
    clocks = None
    if rank == 0:
        t_end = time.time() + 0.4
        while time.time() < t_end:
            run_steps(T)
        torch.cuda.synchronize()
        clocks = sampler.stop()
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())

    # ---- per-kernel timing (roofline): the stepping kernel of one rollout launch, CUDA events on its stream ----
    kr, kf = [], []
    for t in range(24):
        a, b = env.rollout_timed(actions, obs, rew, done, dirs)
        if t >= 4 and a > 0:
            kr.append(a); kf.append(b)          # b = 0 for the launches that do not refill
    k_roll_ms = sum(kr) / len(kr) if kr else 0.0
    k_refill_ms = sum(kf) / len(kf) if kf else 0.0      # amortised per launch
    # the per-step entry point (policy in the loop): bb_pool_step on device buffers, one launch per step
    Ks = min(K, 600)
    for t in range(20):
        env.step(actions[t % T], obs[t % T], rew[t % T], done[t % T], dirs[t % T])
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for t in range(Ks):
        env.step(actions[t % T], obs[t % T], rew[t % T], done[t % T], dirs[t % T])
    s1.record()
    barrier()
    per_step_ms = s0.elapsed_time(s1) / Ks

    # ---- end to end through the host-buffer call ----------------------------------------
    h_act = np.random.RandomState(7 + rank).randint(0, 7, (64, n)).astype(np.int8)
    # page-locked host buffers (what babyai_b200.ParallelEnv hands to bb_pool_step_host)
    pins = [torch.zeros((n, 7, 7, 3), dtype=torch.uint8).pin_memory(), torch.zeros(n, dtype=torch.float32).pin_memory(),
            torch.zeros(n, dtype=torch.uint8).pin_memory(), torch.zeros(n, dtype=torch.int8).pin_memory()]
    h_obs, h_rew, h_done, h_dir = [t.numpy() for t in pins]
    Ke = min(K, 400)
    for k in range(5):
        env.step_host(h_act[k], h_obs, h_rew, h_done, h_dir)
    barrier()
    t0 = time.perf_counter()
    for k in range(Ke):
        env.step_host(h_act[k % 64], h_obs, h_rew, h_done, h_dir)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())

    # ---- the only collective on this path: all-gather of the counters --------------------
    cnt = gather_counters(env.counters(), device=dev)

    # ---- the learner-facing path with observations resident in HBM (babyai_b200.learner; informational) ----------
    # per step: actions host->device, the step kernel writing into a fresh observation tensor, ObssPreprocessor handing
    # the model image float[N,7,7,3] + instr long[N,L] on the device, reward/done device->host.  Measured on this rank.
    learner = {}
    for key, fused_io in (('tensor_copies', False), ('fused_io', True)):
        # tensor_copies: bb_pool_step + torch copies of actions / reward / done; fused_io: bb_pool_step_learner (those three
        # over mapped page-locked memory inside the step call)
        try:
            from babyai_b200 import make_envs
            from babyai_b200.learner import DeviceParallelEnv, ObssPreprocessor
            denv = DeviceParallelEnv(make_envs(args.level, n), pool=env, fused_io=fused_io)
            pre = ObssPreprocessor(trim=False)
            ob = denv.reset()
            Kl = min(K, 300)
            for k in range(5):
                pre(ob, device=dev)
                ob, _r, _d, _i = denv.step(h_act[k])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(Kl):
                pre(ob, device=dev)
                ob, _r, _d, _i = denv.step(h_act[k % 64])
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            learner[key] = {'value_per_gpu': n * Kl / dt, 'unit': 'env-steps/s', 'steps': Kl, 'h2d_bytes_per_step': n,
                            'd2h_bytes_per_step': n * 5}
        except Exception as ex:          # informational leg: never lose the bench line over it
            learner[key] = {'error': repr(ex)[:300]}
    learner['api'] = 'DeviceParallelEnv.step + ObssPreprocessor (observations stay in HBM)'

    # ---- experimental kernel, in a child process with its own CUDA context (single-GPU runs only; informational) ----
    probe = None
    if rank == 0 and world == 1 and not args.no_probe:
        probe = {}
        # the bench workload, and BASELINE config 5's per-GPU share (multi-room: 22x22 staging, k_gen beside the kernel)
        # (key, level, envs, extra environment): the bench workload with the fused generator warp and with refill passes
        # instead (two separate children: whatever happens to one leaves the other's numbers), and BASELINE config 5's
        # per-GPU share (multi-room: 22x22 staging, k_gen beside the kernel)
        jobs = [(args.level, args.level, n, {}), (args.level + ' BB_GEN_FUSED=0', args.level, n, {'BB_GEN_FUSED': '0'})]
        if args.level == LEVEL and n == N_ENVS:
            jobs.append(('BossLevel', 'BossLevel', 32768, {}))
        for key, lv, ne, extra in jobs:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), '--lanes2-probe', '--level', lv, '--envs', str(ne)],
                                   capture_output=True, text=True, timeout=100, env=dict(os.environ, **extra))
                lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
                probe[key] = json.loads(lines[-1]) if lines else {'error': (r.stderr or 'no output')[-400:], 'returncode': r.returncode}
            except subprocess.TimeoutExpired:
                probe[key] = {'error': 'timed out after 100 s (child killed); remaining probes skipped'}
                break                                        # keep the whole bench run within minutes
            except Exception as ex:
                probe[key] = {'error': repr(ex)[:300]}

    if rank == 0:
        peak, peak_src = hbm_peak()
        value = world * n * K / (ms * 1e-3)
        achieved = ALGO_BYTES_PER_STEP * n * T / (k_roll_ms * 1e-3) / 1e9 if k_roll_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get('k_rollout_dram_bytes_per_launch')
            except Exception:
                traffic = None
        out = {
            'metric': METRIC, 'value': value, 'unit': 'env-steps/s', 'n_gpus': world, 'steps': K, 'warmup': max(W, 3),
            'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u8', 'data': 'synthetic',
            'config': {'workload': 'BabyAI-%s-v0, %d envs/GPU, uniform random actions (int8, resident in HBM), '
                                   'ParallelEnv auto-reset, in-kernel verifier' % (args.level, n),
                       'envs_per_gpu': n, 'rollout_chunk': T,
                       'l2_policy': 'obs written to a [%d, %d, 147] buffer (%.0f MB) larger than L2' % (T, n, T * n * 147 / 1e6),
                       'execution': 'bb_pool_rollout: ONE persistent kernel per %d steps (k_rollout: per CTA two stepping warps with '
                                    'the env state resident in shared memory + one generator warp that refills the level rings of the '
                                    "CTA's envs; refill_ms_per_launch > 0 only when BB_GEN_FUSED=0 or for rooms smaller than 6x6)" % T,
                       'parallelism': 'replicas x%d, counters all-gather only' % world},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'kernel': 'k_rollout', 'kernel_ms': k_roll_ms, 'steps_per_launch': T,
                         'kernel_us_per_step': 1e3 * k_roll_ms / T, 'refill_ms_per_launch': k_refill_ms,
                         'algorithmic_bytes_per_launch': ALGO_BYTES_PER_STEP * n * T, 'peak_source': peak_src},
            'per_step_api': {'value': world * n / (per_step_ms * 1e-3), 'unit': 'env-steps/s', 'ms_per_step': per_step_ms,
                             'api': 'bb_pool_step (one launch per step: k_rollout with T = 1 on single-room levels, k_step8 otherwise; level generation on a side stream), device buffers', 'steps': Ks},
            'e2e': {'value': world * n * Ke / e2e_s, 'unit': 'env-steps/s', 'h2d_bytes_per_step': n,
                    'd2h_bytes_per_step': n * (147 + 4 + 1 + 1), 'steps': Ke,
                    'api': 'bb_pool_step_host, page-locked host buffers'},
            'learner_path': learner,
            'experimental': {'k_rollout2': probe},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'counters': cnt,
        }
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, k = cpu_port(12.0, n, threads)
            out['cpu_baseline'] = {'value': v, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'port',
                                   'sample': '%d steps x %d envs of %s, oracle C port, %d host threads' % (k, n, args.level, threads)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
