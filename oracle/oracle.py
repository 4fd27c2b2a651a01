"""ctypes binding of the C oracle (oracle/babyai_oracle.c) + the table that maps
reference level names to the oracle's LevelSpec.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by babyai_b200/.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import build as _build  # noqa: E402

KIND_REDBALL, KIND_OBJ, KIND_LEVELGEN, KIND_IMPUNLOCK, KIND_UNLOCK = 0, 1, 2, 3, 4
I_GOTO, I_PICKUP, I_OPEN, I_PUTNEXT = 0, 1, 2, 3
K_ACTION, K_AND, K_SEQ = 0, 1, 2


class LevelSpec(C.Structure):
    _fields_ = [
        ('kind', C.c_int32), ('room_size', C.c_int32), ('num_rows', C.c_int32), ('num_cols', C.c_int32),
        ('num_dists', C.c_int32), ('instr', C.c_int32), ('doors_open', C.c_int32), ('grey_dists', C.c_int32),
        ('locked_room_prob', C.c_double),
        ('locations', C.c_int32), ('unblocking', C.c_int32), ('implicit_unlock', C.c_int32),
        ('n_action_kinds', C.c_int32), ('action_kinds', C.c_int32 * 4),
        ('n_instr_kinds', C.c_int32), ('instr_kinds', C.c_int32 * 3),
        ('all_unique', C.c_int32), ('require_unreachable', C.c_int32),
    ]


def _redball(num_dists=7, grey=0):                      # iclr19_levels.py:10-72
    return dict(kind=KIND_REDBALL, room_size=8, num_rows=1, num_cols=1, num_dists=num_dists, grey_dists=grey)


def _obj(room_size=8, rows=1, cols=1, num_dists=8, instr=I_GOTO, doors_open=0, all_unique=0, require_unreachable=0):
    # iclr19_levels.py:75-184, :187-221, :224-301, :360-415, :477-491
    return dict(kind=KIND_OBJ, room_size=room_size, num_rows=rows, num_cols=cols, num_dists=num_dists,
                instr=instr, doors_open=doors_open, all_unique=all_unique, require_unreachable=require_unreachable)


def _levelgen(room_size=8, rows=3, cols=3, num_dists=18, locked_room_prob=0.5, locations=1, unblocking=1,
              implicit_unlock=1, action_kinds=(I_GOTO, I_PICKUP, I_OPEN, I_PUTNEXT),
              instr_kinds=(K_ACTION, K_AND, K_SEQ)):   # levelgen.py:262-291
    return dict(kind=KIND_LEVELGEN, room_size=room_size, num_rows=rows, num_cols=cols, num_dists=num_dists,
                locked_room_prob=locked_room_prob, locations=locations, unblocking=unblocking,
                implicit_unlock=implicit_unlock, action_kinds=tuple(action_kinds), instr_kinds=tuple(instr_kinds))


# reference level name -> parameters (babyai/levels/iclr19_levels.py, line cited per family above)
LEVELS = {
    'GoToRedBallGrey': _redball(7, 1),
    'GoToRedBall': _redball(7),
    'GoToRedBallNoDists': _redball(0),
    'GoToObj': _obj(8, num_dists=1), 'GoToObjS4': _obj(4, num_dists=1), 'GoToObjS6': _obj(6, num_dists=1),
    'GoToLocal': _obj(8, num_dists=8),
    'GoToLocalS5N2': _obj(5, num_dists=2), 'GoToLocalS6N2': _obj(6, num_dists=2), 'GoToLocalS6N3': _obj(6, num_dists=3),
    'GoToLocalS6N4': _obj(6, num_dists=4), 'GoToLocalS7N4': _obj(7, num_dists=4), 'GoToLocalS7N5': _obj(7, num_dists=5),
    'GoToLocalS8N2': _obj(8, num_dists=2), 'GoToLocalS8N3': _obj(8, num_dists=3), 'GoToLocalS8N4': _obj(8, num_dists=4),
    'GoToLocalS8N5': _obj(8, num_dists=5), 'GoToLocalS8N6': _obj(8, num_dists=6), 'GoToLocalS8N7': _obj(8, num_dists=7),
    'GoTo': _obj(8, 3, 3, 18), 'GoToOpen': _obj(8, 3, 3, 18, doors_open=1),
    'GoToObjMaze': _obj(8, 3, 3, 1), 'GoToObjMazeOpen': _obj(8, 3, 3, 1, doors_open=1),
    'GoToObjMazeS4R2': _obj(4, 2, 2, 1), 'GoToObjMazeS4': _obj(4, 3, 3, 1), 'GoToObjMazeS5': _obj(5, 3, 3, 1),
    'GoToObjMazeS6': _obj(6, 3, 3, 1), 'GoToObjMazeS7': _obj(7, 3, 3, 1),
    'Pickup': _obj(8, 3, 3, 18, instr=I_PICKUP),
    'UnblockPickup': _obj(8, 3, 3, 20, instr=I_PICKUP, require_unreachable=1),      # :374-391
    'Open': _obj(8, 3, 3, 18, instr=I_OPEN),                                        # :394-415
    'PutNext': _obj(8, 3, 3, 18, instr=I_PUTNEXT),                                  # :477-491
    'PutNextLocal': _obj(8, 1, 1, 8, instr=I_PUTNEXT, all_unique=1),                # :187-211
    'PutNextLocalS5N3': _obj(5, 1, 1, 3, instr=I_PUTNEXT, all_unique=1),
    'PutNextLocalS6N4': _obj(6, 1, 1, 4, instr=I_PUTNEXT, all_unique=1),
    'Unlock': dict(kind=KIND_UNLOCK, room_size=8, num_rows=3, num_cols=3, num_dists=3),          # :418-474
    'GoToImpUnlock': dict(kind=KIND_IMPUNLOCK, room_size=8, num_rows=3, num_cols=3, num_dists=2),   # :304-355
    'PickupLoc': _levelgen(rows=1, cols=1, num_dists=8, locked_room_prob=0, locations=1, unblocking=0,
                           action_kinds=(I_PICKUP,), instr_kinds=(K_ACTION,)),      # :494-515
    'GoToSeq': _levelgen(action_kinds=(I_GOTO,), locked_room_prob=0, locations=0, unblocking=0),   # :518-546
    'GoToSeqS5R2': _levelgen(5, 2, 2, 4, action_kinds=(I_GOTO,), locked_room_prob=0, locations=0, unblocking=0),
    'Synth': _levelgen(instr_kinds=(K_ACTION,), locations=0, unblocking=1, implicit_unlock=0),       # :554-583
    'SynthS5R2': _levelgen(5, 2, 2, 7, instr_kinds=(K_ACTION,), locations=0, unblocking=1, implicit_unlock=0),
    'SynthLoc': _levelgen(instr_kinds=(K_ACTION,), locations=1, unblocking=1, implicit_unlock=0),    # :597-614
    'SynthSeq': _levelgen(locations=1, unblocking=1, implicit_unlock=0),                              # :617-633
    'MiniBossLevel': _levelgen(5, 2, 2, 7, locked_room_prob=0.25),                                    # :636-645
    'BossLevel': _levelgen(),                                                                          # :648-652
    'BossLevelNoUnlock': _levelgen(locked_room_prob=0, implicit_unlock=0),                            # :655-661
}


def make_spec(name):
    d = LEVELS[name]
    s = LevelSpec()
    for k, v in d.items():
        if k == 'action_kinds':
            s.n_action_kinds = len(v)
            for i, x in enumerate(v):
                s.action_kinds[i] = x
        elif k == 'instr_kinds':
            s.n_instr_kinds = len(v)
            for i, x in enumerate(v):
                s.instr_kinds[i] = x
        else:
            setattr(s, k, v)
    return s


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        L = C.CDLL(path)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(LevelSpec), C.c_int]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_seed.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_step.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int]
        L.oracle_step_mt.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int, C.c_int]
        L.oracle_mission.restype = C.c_char_p
        L.oracle_mission.argtypes = [C.c_void_p, C.c_int]
        L.oracle_get_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_width.argtypes = [C.c_void_p]
        L.oracle_height.argtypes = [C.c_void_p]
        L.oracle_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OraclePool:
    """N independent oracle environments of one level (host memory, numpy I/O)."""

    def __init__(self, level, n, seeds=None):
        self.L = lib()
        self.level = level
        self.n = n
        self.spec = make_spec(level)
        self.h = self.L.oracle_create(C.byref(self.spec), n)
        self.width = self.L.oracle_width(self.h)
        self.height = self.L.oracle_height(self.h)
        self.obs = np.zeros((n, 7, 7, 3), np.uint8)
        self.reward = np.zeros(n, np.float32)
        self.done = np.zeros(n, np.uint8)
        self.direction = np.zeros(n, np.int8)
        if seeds is not None:
            self.seed(seeds)

    def __del__(self):
        try:
            self.L.oracle_destroy(self.h)
        except Exception:
            pass

    def seed(self, seeds):
        s = np.ascontiguousarray(seeds, dtype=np.uint64)
        assert s.shape == (self.n,)
        self.L.oracle_seed(self.h, _p(s))

    def reset(self):
        self.L.oracle_reset(self.h, _p(self.obs), _p(self.direction))
        return self.obs

    def step(self, actions, autoreset=True, nthreads=1):
        a = np.ascontiguousarray(actions, dtype=np.int8)
        assert a.shape == (self.n,)
        self.L.oracle_step_mt(self.h, _p(a), _p(self.obs), _p(self.reward), _p(self.done), _p(self.direction),
                              int(autoreset), int(nthreads))
        return self.obs, self.reward, self.done

    def mission(self, i):
        return self.L.oracle_mission(self.h, i).decode()

    def state(self, i):
        grid = np.zeros((self.height, self.width), np.uint8)
        info = np.zeros(8, np.int32)
        self.L.oracle_get_state(self.h, i, _p(grid), _p(info))
        return grid, dict(agent_x=int(info[0]), agent_y=int(info[1]), agent_dir=int(info[2]), carrying=int(info[3]),
                          step_count=int(info[4]), max_steps=int(info[5]), draws=int(info[6]), attempts=int(info[7]))


def philox(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    o = np.zeros(4, np.uint32)
    lib().oracle_philox(_p(c), _p(k), _p(o))
    return tuple(int(x) for x in o)
