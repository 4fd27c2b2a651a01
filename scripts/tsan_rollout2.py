"""k_rollout2 on threads under ThreadSanitizer: every lane is an OS thread (tests/hostemu/simt_rollout2.cpp), so a missing
__syncwarp / __syncthreads in csrc/rollout2.cuh or csrc/gen_round.cuh is a data race TSan reports.  Expected output: exactly
one report -- the `s_done` flag the generator warp polls with a plain volatile read while the stepping warps count
themselves in with an atomic add (deliberate, as on the GPU).

    g++ -O1 -g -std=c++20 -pthread -fsanitize=thread -fno-strict-aliasing -ffp-contract=off -shared -fPIC \
        tests/hostemu/simt_rollout2.cpp -o /tmp/libsimt_tsan.so
    TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" LD_PRELOAD=$(gcc -print-file-name=libtsan.so) python scripts/tsan_rollout2.py
"""
import sys, ctypes as C, numpy as np
sys.path[:0]=['/root/repo','/root/repo/tests','/root/repo/tests/hostemu']
from babyai_b200.levels import level_spec
L = C.CDLL('/tmp/libsimt_tsan.so')
L.r2_create.restype = C.c_void_p
L.r2_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
L.r2_rollout.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5
L.r2_rollout_fused.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
def p(a): return a.ctypes.data_as(C.c_void_p)
for level, fused in (('GoToLocal', False), ('BossLevel', False), ('GoToLocal', True)):
    n, T = 40, 12
    spec = level_spec(level)
    seeds = np.arange(n, dtype=np.uint64) + 5
    h = L.r2_create(C.byref(spec), n, 2 * T + 8, p(seeds), 0)
    rng = np.random.RandomState(0)
    for rep in range(3):
        a = rng.randint(0, 7, (T, n)).astype(np.int8)
        obs, rew = np.zeros((T, n, 147), np.uint8), np.zeros((T, n), np.float32)
        done, dirs, cnt = np.zeros((T, n), np.uint8), np.zeros((T, n), np.int8), np.zeros(4, np.int64)
        if fused: L.r2_rollout_fused(h, p(a), T, 2, 8, p(obs), p(rew), p(done), p(dirs), p(cnt))
        else: L.r2_rollout(h, p(a), T, p(obs), p(rew), p(done), p(dirs), p(cnt))
    print(level, fused, 'done', cnt, flush=True)
