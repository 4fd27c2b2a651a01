"""Level table: reference level name -> constructor arguments of the reference
level class, as a `bb_level_spec` (include/babyai_b200.h).

Each entry cites the class in /root/reference/babyai/levels/iclr19_levels.py it
mirrors.  Five generator families cover all 47 ICLR-19 levels:
  REDBALL  : Level_GoToRedBall* (:10-72)
  OBJ      : place_agent, connect_all, add_distractors, check_objs_reachable,
             pick one (or a door, or two), GoTo / Pickup / Open / PutNext it -- Level_GoToObj* (:75-102),
             Level_GoToLocal* (:105-184), Level_PutNextLocal* (:187-221), Level_GoTo* (:224-301),
             Level_Pickup (:360-371), Level_UnblockPickup (:374-391), Level_Open (:394-415), Level_PutNext (:477-491)
  IMPUNLOCK: Level_GoToImpUnlock (:304-355)
  UNLOCK   : Level_Unlock (:418-474) -- 37 objects: only the doors are table objects, the rest are untracked cell bytes
  LEVELGEN : levelgen.py:256-460 LevelGen -- PickupLoc (:494), GoToSeq (:518), Synth*
             (:554-633), MiniBossLevel (:636), BossLevel (:648), BossLevelNoUnlock (:655)
"""
import ctypes as C
import os

KIND_REDBALL, KIND_OBJ, KIND_LEVELGEN, KIND_IMPUNLOCK, KIND_UNLOCK, KIND_BONUS = 0, 1, 2, 3, 4, 5
I_GOTO, I_PICKUP, I_OPEN, I_PUTNEXT = 0, 1, 2, 3
K_ACTION, K_AND, K_SEQ = 0, 1, 2


class LevelSpec(C.Structure):
    """struct bb_level_spec"""
    _fields_ = [
        ('kind', C.c_int32), ('room_size', C.c_int32), ('num_rows', C.c_int32), ('num_cols', C.c_int32),
        ('num_dists', C.c_int32), ('instr', C.c_int32), ('doors_open', C.c_int32), ('grey_dists', C.c_int32),
        ('locked_room_prob', C.c_double),
        ('locations', C.c_int32), ('unblocking', C.c_int32), ('implicit_unlock', C.c_int32),
        ('n_action_kinds', C.c_int32), ('action_kinds', C.c_int32 * 4),
        ('n_instr_kinds', C.c_int32), ('instr_kinds', C.c_int32 * 3),
        ('all_unique', C.c_int32), ('require_unreachable', C.c_int32),
        ('strict_mask', C.c_int32), ('done_actions', C.c_int32),
        ('bonus', C.c_int32), ('bonus_a', C.c_int32), ('bonus_b', C.c_int32),
    ]


def _spec(kind, room_size=8, num_rows=1, num_cols=1, num_dists=0, instr=I_GOTO, doors_open=0, grey_dists=0,
          locked_room_prob=0.0, locations=0, unblocking=0, implicit_unlock=1, action_kinds=(), instr_kinds=(),
          all_unique=0, require_unreachable=0, strict_mask=0, bonus=0, bonus_a=0, bonus_b=0):
    s = LevelSpec()
    s.kind, s.room_size, s.num_rows, s.num_cols, s.num_dists = kind, room_size, num_rows, num_cols, num_dists
    s.instr, s.doors_open, s.grey_dists = instr, doors_open, grey_dists
    s.locked_room_prob = locked_room_prob
    s.locations, s.unblocking, s.implicit_unlock = locations, unblocking, implicit_unlock
    s.n_action_kinds = len(action_kinds)
    for i, a in enumerate(action_kinds):
        s.action_kinds[i] = a
    s.n_instr_kinds = len(instr_kinds)
    for i, a in enumerate(instr_kinds):
        s.instr_kinds[i] = a
    s.all_unique, s.require_unreachable = all_unique, require_unreachable
    s.strict_mask = strict_mask
    s.bonus, s.bonus_a, s.bonus_b = bonus, bonus_a, bonus_b
    # verifier.use_done_actions is read from the environment when babyai.levels.verifier is imported (verifier.py:15-17)
    s.done_actions = 1 if os.environ.get('BABYAI_DONE_ACTIONS', False) else 0
    return s


def redball(num_dists=7, grey=0):
    return _spec(KIND_REDBALL, 8, 1, 1, num_dists, grey_dists=grey)


def obj_level(room_size=8, num_rows=1, num_cols=1, num_dists=8, instr=I_GOTO, doors_open=0, all_unique=0, require_unreachable=0):
    return _spec(KIND_OBJ, room_size, num_rows, num_cols, num_dists, instr=instr, doors_open=doors_open,
                 all_unique=all_unique, require_unreachable=require_unreachable)


ALL_ACTIONS = (I_GOTO, I_PICKUP, I_OPEN, I_PUTNEXT)
ALL_INSTRS = (K_ACTION, K_AND, K_SEQ)


def levelgen(room_size=8, num_rows=3, num_cols=3, num_dists=18, locked_room_prob=0.5, locations=1, unblocking=1,
             implicit_unlock=1, action_kinds=ALL_ACTIONS, instr_kinds=ALL_INSTRS):
    """LevelGen.__init__ defaults, levelgen.py:262-275"""
    return _spec(KIND_LEVELGEN, room_size, num_rows, num_cols, num_dists, locked_room_prob=locked_room_prob,
                 locations=locations, unblocking=unblocking, implicit_unlock=implicit_unlock,
                 action_kinds=action_kinds, instr_kinds=instr_kinds)


T_KEY, T_BALL, T_BOX = 5, 6, 7
COLOR_IDX = {'red': 0, 'green': 1, 'blue': 2, 'purple': 3, 'yellow': 4, 'grey': 5}


def bonus(family, room_size=8, num_rows=3, num_cols=3, a=0, b=0, num_dists=0, strict_mask=0):
    """babyai/levels/bonus_levels.py: `family` numbers the gen_mission (include/babyai_b200.h, BB_KIND_BONUS)"""
    return _spec(KIND_BONUS, room_size, num_rows, num_cols, num_dists, strict_mask=strict_mask, bonus=family, bonus_a=a, bonus_b=b)


LEVELS = {
    'GoToRedBallGrey': lambda: redball(7, 1),
    'GoToRedBall': lambda: redball(7),
    'GoToRedBallNoDists': lambda: redball(0),
    'GoToObj': lambda: obj_level(8, num_dists=1),
    'GoToObjS4': lambda: obj_level(4, num_dists=1),
    'GoToObjS6': lambda: obj_level(6, num_dists=1),
    'GoToLocal': lambda: obj_level(8, num_dists=8),
    'GoToLocalS5N2': lambda: obj_level(5, num_dists=2),
    'GoToLocalS6N2': lambda: obj_level(6, num_dists=2),
    'GoToLocalS6N3': lambda: obj_level(6, num_dists=3),
    'GoToLocalS6N4': lambda: obj_level(6, num_dists=4),
    'GoToLocalS7N4': lambda: obj_level(7, num_dists=4),
    'GoToLocalS7N5': lambda: obj_level(7, num_dists=5),
    'GoToLocalS8N2': lambda: obj_level(8, num_dists=2),
    'GoToLocalS8N3': lambda: obj_level(8, num_dists=3),
    'GoToLocalS8N4': lambda: obj_level(8, num_dists=4),
    'GoToLocalS8N5': lambda: obj_level(8, num_dists=5),
    'GoToLocalS8N6': lambda: obj_level(8, num_dists=6),
    'GoToLocalS8N7': lambda: obj_level(8, num_dists=7),
    'GoTo': lambda: obj_level(8, 3, 3, 18),
    'GoToOpen': lambda: obj_level(8, 3, 3, 18, doors_open=1),
    'GoToObjMaze': lambda: obj_level(8, 3, 3, 1),
    'GoToObjMazeOpen': lambda: obj_level(8, 3, 3, 1, doors_open=1),
    'GoToObjMazeS4R2': lambda: obj_level(4, 2, 2, 1),
    'GoToObjMazeS4': lambda: obj_level(4, 3, 3, 1),
    'GoToObjMazeS5': lambda: obj_level(5, 3, 3, 1),
    'GoToObjMazeS6': lambda: obj_level(6, 3, 3, 1),
    'GoToObjMazeS7': lambda: obj_level(7, 3, 3, 1),
    'Pickup': lambda: obj_level(8, 3, 3, 18, instr=I_PICKUP),
    'UnblockPickup': lambda: obj_level(8, 3, 3, 20, instr=I_PICKUP, require_unreachable=1),
    'Open': lambda: obj_level(8, 3, 3, 18, instr=I_OPEN),
    'PutNext': lambda: obj_level(8, 3, 3, 18, instr=I_PUTNEXT),
    'PutNextLocal': lambda: obj_level(8, 1, 1, 8, instr=I_PUTNEXT, all_unique=1),
    'PutNextLocalS5N3': lambda: obj_level(5, 1, 1, 3, instr=I_PUTNEXT, all_unique=1),
    'PutNextLocalS6N4': lambda: obj_level(6, 1, 1, 4, instr=I_PUTNEXT, all_unique=1),
    'GoToImpUnlock': lambda: _spec(KIND_IMPUNLOCK, 8, 3, 3, num_dists=2),
    'Unlock': lambda: _spec(KIND_UNLOCK, 8, 3, 3, num_dists=3),
    'PickupLoc': lambda: levelgen(num_rows=1, num_cols=1, num_dists=8, locked_room_prob=0, locations=1, unblocking=0,
                                  action_kinds=(I_PICKUP,), instr_kinds=(K_ACTION,)),
    'GoToSeq': lambda: levelgen(action_kinds=(I_GOTO,), locked_room_prob=0, locations=0, unblocking=0),
    'GoToSeqS5R2': lambda: levelgen(5, 2, 2, 4, action_kinds=(I_GOTO,), locked_room_prob=0, locations=0, unblocking=0),
    'Synth': lambda: levelgen(instr_kinds=(K_ACTION,), locations=0, unblocking=1, implicit_unlock=0),
    'SynthS5R2': lambda: levelgen(5, 2, 2, 7, instr_kinds=(K_ACTION,), locations=0, unblocking=1, implicit_unlock=0),
    'SynthLoc': lambda: levelgen(instr_kinds=(K_ACTION,), locations=1, unblocking=1, implicit_unlock=0),
    'SynthSeq': lambda: levelgen(locations=1, unblocking=1, implicit_unlock=0),
    'MiniBossLevel': lambda: levelgen(5, 2, 2, 7, locked_room_prob=0.25),
    'BossLevel': lambda: levelgen(),
    'BossLevelNoUnlock': lambda: levelgen(locked_room_prob=0, implicit_unlock=0),
    # ---- bonus_levels.py (class Level_<name>) ----
    'GoToRedBlueBall': lambda: bonus(1, 8, 1, 1, num_dists=7),
    'OpenRedDoor': lambda: bonus(2, 5, 1, 2),
    'OpenDoor': lambda: bonus(3),
    'OpenDoorDebug': lambda: bonus(3, strict_mask=1),
    'OpenDoorColor': lambda: bonus(3, a=1),
    'OpenDoorLoc': lambda: bonus(3, a=2),
    'GoToDoor': lambda: bonus(4, 7),
    'GoToObjDoor': lambda: bonus(5, 8),
    'ActionObjDoor': lambda: bonus(6, 7),
    'UnlockLocal': lambda: bonus(7),
    'UnlockLocalDist': lambda: bonus(7, a=1),
    'KeyInBox': lambda: bonus(8),
    'UnlockPickup': lambda: bonus(9, 6, 1, 2),
    'UnlockPickupDist': lambda: bonus(9, 6, 1, 2, a=1),
    'BlockedUnlockPickup': lambda: bonus(10, 6, 1, 2),
    'UnlockToUnlock': lambda: bonus(11, 6, 1, 3),
    'PickupDist': lambda: bonus(12, 7, 1, 1),
    'PickupDistDebug': lambda: bonus(12, 7, 1, 1, strict_mask=1),
    'PickupAbove': lambda: bonus(13, 6),
    'OpenTwoDoors': lambda: bonus(14, 6),
    'OpenTwoDoorsDebug': lambda: bonus(14, 6, strict_mask=1),
    'OpenRedBlueDoors': lambda: bonus(14, 6, a=1 + COLOR_IDX['red'], b=1 + COLOR_IDX['blue']),
    'OpenRedBlueDoorsDebug': lambda: bonus(14, 6, a=1 + COLOR_IDX['red'], b=1 + COLOR_IDX['blue'], strict_mask=1),
    'FindObjS5': lambda: bonus(15, 5),
    'FindObjS6': lambda: bonus(15, 6),
    'FindObjS7': lambda: bonus(15, 7),
    'KeyCorridorS3R1': lambda: bonus(16, 3, 1, 3, a=T_BALL),
    'KeyCorridorS3R2': lambda: bonus(16, 3, 2, 3, a=T_BALL),
    'KeyCorridorS3R3': lambda: bonus(16, 3, 3, 3, a=T_BALL),
    'KeyCorridorS4R3': lambda: bonus(16, 4, 3, 3, a=T_BALL),
    'KeyCorridorS5R3': lambda: bonus(16, 5, 3, 3, a=T_BALL),
    'KeyCorridorS6R3': lambda: bonus(16, 6, 3, 3, a=T_BALL),
    '1RoomS8': lambda: bonus(17, 8, 1, 1),
    '1RoomS12': lambda: bonus(17, 12, 1, 1),
    '1RoomS16': lambda: bonus(17, 16, 1, 1),
    '1RoomS20': lambda: bonus(17, 20, 1, 1),
    'PutNextS4N1': lambda: bonus(18, 4, 1, 2, a=1),
    'PutNextS5N1': lambda: bonus(18, 5, 1, 2, a=1),
    'PutNextS5N2': lambda: bonus(18, 5, 1, 2, a=2),
    'PutNextS6N3': lambda: bonus(18, 6, 1, 2, a=3),
    'PutNextS7N4': lambda: bonus(18, 7, 1, 2, a=4),
    'PutNextS5N2Carrying': lambda: bonus(18, 5, 1, 2, a=2, b=1),
    'PutNextS6N3Carrying': lambda: bonus(18, 6, 1, 2, a=3, b=1),
    'PutNextS7N4Carrying': lambda: bonus(18, 7, 1, 2, a=4, b=1),
    'MoveTwoAcrossS5N2': lambda: bonus(19, 5, 1, 2, a=2),
    'MoveTwoAcrossS8N9': lambda: bonus(19, 8, 1, 2, a=9),
    'OpenDoorsOrderN2': lambda: bonus(20, 6, a=2),
    'OpenDoorsOrderN4': lambda: bonus(20, 6, a=4),
    'OpenDoorsOrderN2Debug': lambda: bonus(20, 6, a=2, strict_mask=5),
    'OpenDoorsOrderN4Debug': lambda: bonus(20, 6, a=4, strict_mask=5),
}
ICLR19_LEVELS = [k for k in LEVELS if LEVELS[k]().kind != KIND_BONUS]
BONUS_LEVELS = [k for k in LEVELS if LEVELS[k]().kind == KIND_BONUS]


def level_spec(name):
    """'GoToLocal' or the gym id 'BabyAI-GoToLocal-v0' (levelgen.py:481)."""
    if name.startswith('BabyAI-') and name.endswith('-v0'):
        name = name[len('BabyAI-'):-len('-v0')]
    if name not in LEVELS:
        raise KeyError('level %r is not supported by the B200 pool (supported: %s)' % (name, ', '.join(sorted(LEVELS))))
    return LEVELS[name]()


# fixed vocabulary of the baby language; index = token id (0 = padding).
# Matches bb_vocab_word() / the W_* enum in csrc/env_logic.cuh.
VOCAB = ['', 'go', 'to', 'pick', 'up', 'open', 'put', 'next', 'the', 'a', 'object',
         'red', 'green', 'blue', 'purple', 'yellow', 'grey', 'box', 'ball', 'key', 'door',
         'in', 'front', 'of', 'you', 'behind', 'on', 'your', 'left', 'right', 'then', 'after', 'and']


def detokenize(tokens):
    """int16 token row -> the reference's mission string (Instr.surface, verifier.py)."""
    words = [VOCAB[int(t)] for t in tokens if int(t) != 0]
    out = []
    for i, w in enumerate(words):
        # BeforeInstr.surface joins with ', then ' (verifier.py:440)
        if w == 'then' and out:
            out[-1] = out[-1] + ','
        out.append(w)
    return ' '.join(out)
