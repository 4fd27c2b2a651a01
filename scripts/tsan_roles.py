"""The kernels' role functions on OS threads under ThreadSanitizer: every lane is an OS thread (tests/hostemu/simt_rollout.cpp), so a
missing __syncwarp / named barrier between two lanes' shared-memory accesses in csrc/rollout_lane.cuh, rollout_cta.cuh, step8.cuh
or gen_round.cuh is a data race TSan reports.  TEST INFRASTRUCTURE (CPU only).

    g++ -O1 -g -std=c++20 -pthread -fsanitize=thread -fno-strict-aliasing -ffp-contract=off -shared -fPIC \\
        tests/hostemu/simt_rollout.cpp -o /tmp/libsimt_tsan.so
    TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" LD_PRELOAD=$(gcc -print-file-name=libtsan.so) python scripts/tsan_roles.py

Expected: reports only for the ONE deliberate race -- the `s_done` word the fused generator warp polls with a plain volatile read
while the stepping warps count themselves in (as on the GPU) -- and for same-value stores of the generation pass between launches
(host `refill`, single-threaded).  Last run: profiles/r02z_tsan_roles.log."""
import ctypes as C
import sys

import numpy as np

sys.path[:0] = ['/root/repo', '/root/repo/tests', '/root/repo/tests/hostemu']
from babyai_b200.levels import level_spec  # noqa: E402

L = C.CDLL(sys.argv[1] if len(sys.argv) > 1 else '/tmp/libsimt_tsan.so')
L.r2_create.restype = C.c_void_p
L.r2_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
L.r2_rollout.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5
L.r2_rollout_cta.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5
L.r2_rollout_fused.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
L.r2_step8.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p]


def p(a):
    return a.ctypes.data_as(C.c_void_p)


for level, kind in (('GoToLocal', 'lane'), ('GoToLocal', 'fused'), ('GoToObjMazeS4R2', 'cta'), ('BossLevel', 'cta'), ('Unlock', 'cta'),
                    ('GoToObjMazeS4R2', 'step8'), ('BossLevel', 'step8')):
    n, T = 40, 12
    spec = level_spec(level)
    seeds = np.arange(n, dtype=np.uint64) + 5
    h = L.r2_create(C.byref(spec), n, 2 * T + 8, p(seeds), 0)
    rng = np.random.RandomState(0)
    cnt = np.zeros(4, np.int64)
    for rep in range(3):
        a = rng.randint(0, 7, (T, n)).astype(np.int8)
        obs, rew = np.zeros((T, n, 147), np.uint8), np.zeros((T, n), np.float32)
        done, dirs = np.zeros((T, n), np.uint8), np.zeros((T, n), np.int8)
        if kind == 'fused':
            L.r2_rollout_fused(h, p(a), T, 2, 8, p(obs), p(rew), p(done), p(dirs), p(cnt))
        elif kind == 'cta':
            L.r2_rollout_cta(h, p(a), T, p(obs), p(rew), p(done), p(dirs), p(cnt))
        elif kind == 'step8':
            for t in range(T):
                L.r2_step8(h, p(a[t]), p(obs[t]), p(rew[t]), p(done[t]), p(dirs[t]), 0, p(cnt))
        else:
            L.r2_rollout(h, p(a), T, p(obs), p(rew), p(done), p(dirs), p(cnt))
    print(level, kind, 'done', cnt, flush=True)
