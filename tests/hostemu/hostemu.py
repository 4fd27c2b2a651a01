"""TEST-ONLY host build of the kernels' per-environment source
(babyai_b200/csrc/env_logic.cuh compiled with g++).  Used by `not gpu` tests to
exercise generation / step / verifier / observation / staging logic in the
container that has no GPU.  Never imported by babyai_b200/."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, '..', '..'))
SRC = os.path.join(HERE, 'hostemu.cpp')
OUT = os.path.join(HERE, 'libhostemu.so')
DEPS = [SRC, os.path.join(ROOT, 'babyai_b200', 'csrc', 'env_logic.cuh'), os.path.join(ROOT, 'babyai_b200', 'csrc', 'level_params.h'), os.path.join(ROOT, 'include', 'babyai_b200.h')]


def build():
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    subprocess.check_call(['g++', '-O1', '-g', '-std=c++17', '-Wall', '-Wno-unknown-pragmas', '-ffp-contract=off',
                           '-shared', '-fPIC', SRC, '-o', OUT])
    return OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.he_create.restype = C.c_void_p
        L.he_create.argtypes = [C.c_void_p, C.c_int]
        for f in ('he_destroy',):
            getattr(L, f).argtypes = [C.c_void_p]
        L.he_set_mode.argtypes = [C.c_void_p, C.c_int]
        L.he_seed.argtypes = [C.c_void_p, C.c_void_p]
        L.he_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.he_step.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.he_tokens.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.he_get_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.he_check_room_obs.argtypes = [C.c_void_p, C.c_int]
        L.he_width.argtypes = [C.c_void_p]
        L.he_height.argtypes = [C.c_void_p]
        L.he_vis_rows.argtypes = [C.c_void_p, C.c_void_p]
        L.he_stage.argtypes = [C.c_void_p, C.c_void_p]
        L.he_stage21.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class HostEmuPool:
    def __init__(self, spec, n, seeds=None, mode=0):
        self.L = lib()
        self.n = n
        self.spec = spec
        self.h = self.L.he_create(C.byref(spec), n)
        self.L.he_set_mode(self.h, mode)
        self.width = self.L.he_width(self.h)
        self.height = self.L.he_height(self.h)
        self.obs = np.zeros((n, 7, 7, 3), np.uint8)
        self.reward = np.zeros(n, np.float32)
        self.done = np.zeros(n, np.uint8)
        self.direction = np.zeros(n, np.int8)
        if seeds is not None:
            self.seed(seeds)

    def __del__(self):
        try:
            self.L.he_destroy(self.h)
        except Exception:
            pass

    def seed(self, seeds):
        s = np.ascontiguousarray(seeds, dtype=np.uint64)
        self.L.he_seed(self.h, _p(s))

    def reset(self):
        self.L.he_reset(self.h, _p(self.obs), _p(self.direction))
        return self.obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int8)
        self.L.he_step(self.h, _p(a), _p(self.obs), _p(self.reward), _p(self.done), _p(self.direction))
        return self.obs, self.reward, self.done

    def tokens(self, i):
        t = np.zeros(72, np.int16)
        self.L.he_tokens(self.h, i, _p(t))
        return t

    def state(self, i):
        grid = np.zeros((self.height, self.width), np.uint8)
        info = np.zeros(8, np.int32)
        self.L.he_get_state(self.h, i, _p(grid), _p(info))
        return grid, dict(agent_x=int(info[0]), agent_y=int(info[1]), agent_dir=int(info[2]), carrying=int(info[3]),
                          step_count=int(info[4]), max_steps=int(info[5]), draws=int(info[6]), attempts=int(info[7]))


# ---- the rollout kernels' roles on OS threads (simt_rollout.cpp) -------------------------------------------------
SRC2 = os.path.join(HERE, 'simt_rollout.cpp')
OUT2 = os.path.join(HERE, 'libsimt_rollout.so')
DEPS2 = [SRC2, os.path.join(ROOT, 'babyai_b200', 'csrc', 'simt.cuh'), os.path.join(ROOT, 'babyai_b200', 'csrc', 'rollout_lane.cuh'), os.path.join(ROOT, 'babyai_b200', 'csrc', 'rollout_cta.cuh'), os.path.join(ROOT, 'babyai_b200', 'csrc', 'step8.cuh'), os.path.join(ROOT, 'babyai_b200', 'csrc', 'gen_round.cuh')] + DEPS[1:]
_lib2 = None


def lib2():
    global _lib2
    if _lib2 is None:
        if not (os.path.exists(OUT2) and all(os.path.getmtime(OUT2) >= os.path.getmtime(d) for d in DEPS2)):
            subprocess.check_call(['g++', '-O1', '-g', '-std=c++20', '-pthread', '-Wall', '-Wno-unknown-pragmas', '-Wno-unused-function',
                                   '-fno-strict-aliasing', '-ffp-contract=off', '-shared', '-fPIC', SRC2, '-o', OUT2])
        L = C.CDLL(OUT2)
        L.r2_create.restype = C.c_void_p
        L.r2_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.r2_destroy.argtypes = [C.c_void_p]
        L.r2_rollout.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.r2_tokens.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.r2_rollout_fused.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
        L.r2_min_ring_level.argtypes = [C.c_void_p]
        L.r2_max_tokens.argtypes = [C.c_void_p]
        L.r2_error_flag.argtypes = [C.c_void_p]
        L.r2_rollout_cta.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.r2_step8.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
        L.r2_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib2 = L
    return _lib2


class RolloutPool:
    """A pool in pool.cu's memory layout (SoA + level rings) stepped by the rollout kernels' role functions, one OS thread per lane."""

    def __init__(self, spec, n, seeds, depth=24, mode=0):
        self.L, self.n = lib2(), n
        s = np.ascontiguousarray(seeds, dtype=np.uint64)
        self.h = self.L.r2_create(C.byref(spec), n, depth, _p(s), mode)

    def __del__(self):
        try:
            self.L.r2_destroy(self.h)
        except Exception:
            pass

    def rollout(self, actions, fused=False, gen_rounds=1 << 20, gen_min_active=0, kernel='lane'):
        """fused=True: the CTA's generator warp refills the rings during the launch (nothing else does);
        kernel='cta': k_rollout_cta's role (rollout_cta.cuh) instead of k_rollout's (rollout_lane.cuh)"""
        a = np.ascontiguousarray(actions, dtype=np.int8)
        T, n = a.shape
        obs, rew = np.zeros((T, n, 7, 7, 3), np.uint8), np.zeros((T, n), np.float32)
        done, dirs, cnt = np.zeros((T, n), np.uint8), np.zeros((T, n), np.int8), np.zeros(4, np.int64)
        if kernel == 'cta':
            self.L.r2_rollout_cta(self.h, _p(a), T, _p(obs), _p(rew), _p(done), _p(dirs), _p(cnt))
        elif fused:
            self.L.r2_rollout_fused(self.h, _p(a), T, gen_rounds, gen_min_active, _p(obs), _p(rew), _p(done), _p(dirs), _p(cnt))
        else:
            self.L.r2_rollout(self.h, _p(a), T, _p(obs), _p(rew), _p(done), _p(dirs), _p(cnt))
        return obs, rew, done, dirs, cnt

    def step8(self, actions, force_reset=0):
        """one bb_pool_step through k_step8's role function (step8.cuh), 128 OS threads per CTA"""
        a = np.ascontiguousarray(actions, dtype=np.int8)
        n = a.shape[0]
        obs, rew = np.zeros((n, 7, 7, 3), np.uint8), np.zeros(n, np.float32)
        done, dirs, cnt = np.zeros(n, np.uint8), np.zeros(n, np.int8), np.zeros(4, np.int64)
        self.L.r2_step8(self.h, _p(a), _p(obs), _p(rew), _p(done), _p(dirs), force_reset, _p(cnt))
        return obs, rew, done, dirs, cnt

    def state(self, i, width, height):
        grid, info = np.zeros((height, width), np.uint8), np.zeros(8, np.int32)
        self.L.r2_state(self.h, i, _p(grid), _p(info))
        return grid, info

    def error_flag(self):
        return self.L.r2_error_flag(self.h)

    def min_ring_level(self):
        return self.L.r2_min_ring_level(self.h)

    def tokens(self, i):
        t = np.zeros(self.L.r2_max_tokens(self.h), np.int16)
        self.L.r2_tokens(self.h, i, _p(t))
        return t
