"""The kernels' per-environment source (babyai_b200/csrc/env_logic.cuh), compiled
for the host by tests/hostemu, against the golden traces and the oracle.  This is
how the CUDA logic is exercised in the GPU-less container; the product itself
never runs on the CPU."""
import itertools

import numpy as np
import pytest

import hostemu
import oracle as orc
from babyai_b200.levels import LEVELS, detokenize, level_spec
from common import GOLDEN_LEVELS, SUCCESS_GOLDENS, compare_pools, replay_golden


def _emu(level, n, seeds, mode=0):
    return hostemu.HostEmuPool(level_spec(level), n, seeds, mode)


@pytest.mark.parametrize('level', GOLDEN_LEVELS)
def test_emu_replays_golden(level):
    replay_golden(level, _emu, lambda p, i: detokenize(p.tokens(i)))


@pytest.mark.parametrize('name', SUCCESS_GOLDENS)
def test_hostemu_replays_success_golden(name):
    replay_golden(name, _emu, lambda p, i: detokenize(p.tokens(i)))


@pytest.mark.parametrize('level', sorted(LEVELS))
def test_emu_matches_oracle_random_actions(level):
    n = 8
    seeds = np.arange(n, dtype=np.uint64) * 7 + 31
    steps = 120 if LEVELS[level]().num_rows > 1 else 200
    compare_pools(orc.OraclePool(level, n, seeds), _emu(level, n, seeds), n, steps,
                  mission_a=lambda p, i: p.mission(i), mission_b=lambda p, i: detokenize(p.tokens(i)))


@pytest.mark.parametrize('level', ['PickupLoc', 'PutNextLocal', 'PutNextLocalS5N3', 'Synth', 'MiniBossLevel', 'Open', 'BossLevel', 'GoToImpUnlock', 'Unlock'])
def test_emu_matches_oracle_interaction_heavy_actions(level):
    """Actions biased towards forward / pickup / drop / toggle so that objects are carried around, boxes opened,
    doors toggled and the obj_poss snapshots go stale and get refreshed (verifier.py:195-202, levelgen.py:53-54)."""
    n = 12
    seeds = np.arange(n, dtype=np.uint64) * 3 + 77
    p = [0.12, 0.12, 0.30, 0.17, 0.14, 0.13, 0.02]
    compare_pools(orc.OraclePool(level, n, seeds), _emu(level, n, seeds), n, 400, act_seed=5, action_p=p,
                  mission_a=lambda q, i: q.mission(i), mission_b=lambda q, i: detokenize(q.tokens(i)))


@pytest.mark.parametrize('level', ['GoToRedBallGrey', 'GoToLocal', 'GoToObjS4', 'GoToLocalS5N2', 'PickupLoc', 'PutNextLocal'])
def test_room_observation_equals_generic(level):
    """Single-room levels use the closed-form visibility (visible <=> inside the grid): it must equal the general
    process_vis path (minigrid.py:1211-1245 restated in vis_rows) and the literal loops for every pose on real levels."""
    n = 24
    e = _emu(level, n, np.arange(n, dtype=np.uint64) + 900)
    e.reset()
    for i in range(n):
        c = e.L.he_check_room_obs(e.h, i)
        interior = (e.width - 2) * (e.height - 2)
        assert c == interior * 4 * 2, (level, i, c)


@pytest.mark.timeout(120)
def test_unsatisfiable_agent_room_is_rejected_not_spun_on():
    """MiniBossLevel seed 698, 57th level: RoomGrid.place_agent's `while True` can never succeed in the room it
    drew (the reference hangs there, tests/test_oracle_vs_reference.py); generation must reject and move on."""
    seeds = np.array([698], dtype=np.uint64)
    e, o = _emu('MiniBossLevel', 1, seeds), orc.OraclePool('MiniBossLevel', 1, seeds)
    for k in range(80):
        assert np.array_equal(e.reset(), np.asarray(o.reset())), k
        assert detokenize(e.tokens(0)) == o.mission(0)


@pytest.mark.timeout(300)
@pytest.mark.parametrize('level', ['MiniBossLevel', 'GoToObjMazeS4', 'SynthS5R2', 'BossLevel', 'GoToLocal', 'PutNextLocal', 'GoToImpUnlock', 'Unlock'])
def test_deep_generation_matches_oracle(level):
    """Generation far down each env's random stream (the GPU pool pre-generates 128 levels per env)."""
    n = 48
    seeds = np.arange(n, dtype=np.uint64) * 7 + 5000
    e, o = _emu(level, n, seeds), orc.OraclePool(level, n, seeds)
    for k in range(400):
        assert np.array_equal(e.reset(), np.asarray(o.reset())), k


def test_freeze_mode_matches_oracle_without_autoreset():
    """ManyEnvs flavour (evaluate.py:72-78): finished envs stop and replay their last result."""
    level, n = 'GoToLocal', 16
    seeds = np.arange(n, dtype=np.uint64) + 900
    o = orc.OraclePool(level, n, seeds)
    e = _emu(level, n, seeds, mode=1)
    assert np.array_equal(o.reset(), e.reset())
    # draws / attempts are comparable here: nothing is pre-generated in freeze mode
    for i in range(n):
        assert o.state(i)[1] == e.state(i)[1]
    rng = np.random.RandomState(3)
    last = [None] * n
    frozen = np.zeros(n, bool)
    for t in range(80):
        act = rng.randint(0, 7, n).astype(np.int8)
        eo, er, ed = [x.copy() for x in e.step(act)]
        for i in range(n):
            if frozen[i]:
                assert (eo[i] == last[i][0]).all() and er[i] == last[i][1] and ed[i]
        oo, orr, od = o.step(act, autoreset=False)
        for i in range(n):
            if not frozen[i]:
                assert (eo[i] == oo[i]).all() and er[i] == orr[i] and ed[i] == od[i]
                if od[i]:
                    frozen[i] = True
                    last[i] = (oo[i].copy(), orr[i])
        # the oracle keeps stepping finished envs; stop comparing those
    assert frozen.any()


def _literal_process_vis(see):
    """gym_minigrid Grid.process_vis, literally (App. A.5), on a see-through table."""
    mask = np.zeros((7, 7), bool)
    mask[3, 6] = True
    for j in reversed(range(7)):
        for i in range(0, 6):
            if not mask[i, j] or not see[i][j]:
                continue
            mask[i + 1, j] = True
            if j > 0:
                mask[i + 1, j - 1] = True
                mask[i, j - 1] = True
        for i in reversed(range(1, 7)):
            if not mask[i, j] or not see[i][j]:
                continue
            mask[i - 1, j] = True
            if j > 0:
                mask[i - 1, j - 1] = True
                mask[i, j - 1] = True
    return mask


def test_vis_rows_bit_trick_equals_literal_loops():
    """Exhaustive over pairs of adjacent rows x random rest + 20k fully random patterns."""
    import ctypes as C
    L = hostemu.lib()
    rng = np.random.RandomState(0)
    pats = []
    for a, b in itertools.product(range(128), range(128)):       # rows 6 and 5 exhaustive, others random
        rows = rng.randint(0, 128, 7)
        rows[6], rows[5] = a, b
        pats.append(rows)
    pats += [rng.randint(0, 128, 7) for _ in range(20000)]
    pats += [np.full(7, 127), np.zeros(7, int)]
    for rows in pats:
        see_rows = np.asarray(rows, np.uint32)
        vis = np.zeros(7, np.uint32)
        L.he_vis_rows(see_rows.ctypes.data_as(C.c_void_p), vis.ctypes.data_as(C.c_void_p))
        see = [[(int(see_rows[j]) >> i) & 1 for j in range(7)] for i in range(7)]
        lit = _literal_process_vis(see)
        got = np.array([[(int(vis[j]) >> i) & 1 for j in range(7)] for i in range(7)], bool)
        assert np.array_equal(lit, got), rows


def test_warp_staging_packs_147_byte_records():
    """32 lanes x 37 words -> the packed 4704-byte tile (byte-exact, every lane misalignment)."""
    import ctypes as C
    L = hostemu.lib()
    rng = np.random.RandomState(1)
    for _ in range(50):
        by = rng.randint(0, 256, (32, 147)).astype(np.uint8)
        w = np.zeros((32, 148), np.uint8)
        w[:, :147] = by
        words = np.ascontiguousarray(w).view(np.uint32).reshape(32, 37)
        tile = np.zeros(4704 + 16, np.uint8)
        L.he_stage(words.ctypes.data_as(C.c_void_p), tile.ctypes.data_as(C.c_void_p))
        assert np.array_equal(tile[:4704], by.reshape(-1))


def test_column_record_staging_packs_21_byte_records():
    """28 view-column records of 21 bytes (4 envs x 7 columns) -> the packed 588-byte warp tile."""
    import ctypes as C
    L = hostemu.lib()
    rng = np.random.RandomState(2)
    for _ in range(50):
        by = rng.randint(0, 256, (28, 21)).astype(np.uint8)
        w = np.zeros((28, 24), np.uint8)
        w[:, :21] = by
        words = np.ascontiguousarray(w).view(np.uint32).reshape(28, 6)
        tile = np.zeros(588 + 16, np.uint8)
        L.he_stage21(words.ctypes.data_as(C.c_void_p), tile.ctypes.data_as(C.c_void_p))
        assert np.array_equal(tile[:588], by.reshape(-1))


@pytest.mark.timeout(600)
@pytest.mark.parametrize('kernel', ['lane', 'pipe'])
@pytest.mark.parametrize('level,n,T', [('GoToLocal', 40, 40), ('PickupLoc', 35, 40), ('GoToObjS4', 16, 40), ('PutNextLocal', 24, 32),
                                        ('GoToObjMazeS4R2', 21, 40), ('BossLevel', 18, 16), ('Unlock', 20, 24)])
def test_rollout_stepping_role_on_threads(level, n, T, kernel):
    """k_rollout's stepping role (babyai_b200/csrc/rollout_lane.cuh: the very function the kernel calls) executed with one OS
    thread per lane and the warp shuffles / votes / barriers as rendezvous: whole rollouts -- step, warp-cooperative level
    swap-in from the ring, observation, tile staging, (bulk) tile stores, state write-back, counters -- must equal the
    per-step path (which is checked against the oracle), ragged last warps included."""
    seeds = np.arange(n, dtype=np.uint64) * 11 + 321
    ref = _emu(level, n, seeds)
    r2 = hostemu.RolloutPool(level_spec(level), n, seeds, depth=max(24, T + 8))
    obs0 = ref.reset().copy()
    rng = np.random.RandomState(17)
    steps = episodes = 0
    for rep in range(5 if level == 'PutNextLocal' else 3):             # PutNextLocal: max_steps = 128
        acts = rng.choice(7, size=(T, n), p=[0.13, 0.13, 0.4, 0.12, 0.08, 0.12, 0.02]).astype(np.int8)
        obs, rew, done, dirs, cnt = r2.rollout(acts, kernel=kernel)
        for t in range(T):
            o, r, d = ref.step(acts[t])
            assert np.array_equal(obs[t], o), (level, rep, t, np.nonzero((obs[t] != o).reshape(n, -1).any(1))[0])
            assert np.array_equal(rew[t].view(np.uint32), r.view(np.uint32)) and np.array_equal(done[t], d), (level, rep, t)
            assert np.array_equal(dirs[t], ref.direction), (level, rep, t)
            episodes += int(d.sum())
        steps += T * n
        assert cnt[0] == steps and cnt[1] == episodes and cnt[3] == 0, (cnt, steps, episodes)
        assert all(np.array_equal(r2.tokens(i)[:8], ref.tokens(i)[:8]) for i in range(n))
    assert episodes > 0 or level in ('BossLevel', 'Unlock')
    del obs0


@pytest.mark.timeout(900)
@pytest.mark.parametrize('level,n,T,mode', [('BossLevel', 40, 16, 0), ('GoTo', 70, 24, 0), ('GoToObjMazeS4R2', 45, 40, 0), ('Unlock', 33, 24, 0),
                                             ('GoToLocal', 50, 40, 0), ('PutNextLocal', 20, 32, 0), ('MiniBossLevel', 64, 40, 0), ('GoToObjMazeS4R2', 37, 40, 1)])
def test_rollout_cta_role_on_threads(level, n, T, mode):
    """k_rollout_cta's role (babyai_b200/csrc/rollout_cta.cuh: the very function the kernel calls) with one OS thread per
    lane, 128 threads per CTA: lane-per-env step phase, CTA-wide swap-in, 4-lanes-per-env observation from the row-major
    grid alone (byte gathers, xor-shuffle exchange of the see-through masks, 21-byte record staging), tile store, and the
    write-back that regenerates the transposed grid -- must equal the per-step path, ragged last CTA included; afterwards
    the hidden state (both grid orientations, pose, carried object, step counter) equals it too."""
    seeds = np.arange(n, dtype=np.uint64) * 13 + 99
    ref = _emu(level, n, seeds, mode=mode)
    r2 = hostemu.RolloutPool(level_spec(level), n, seeds, depth=max(24, T + 8), mode=mode)
    ref.reset()
    rng = np.random.RandomState(29)
    steps = episodes = 0
    for rep in range(4):
        acts = rng.choice(7, size=(T, n), p=[0.13, 0.13, 0.4, 0.12, 0.08, 0.12, 0.02]).astype(np.int8)
        obs, rew, done, dirs, cnt = r2.rollout(acts, kernel='cta')
        for t in range(T):
            o, r, d = ref.step(acts[t])
            assert np.array_equal(obs[t], o), (level, rep, t, np.nonzero((obs[t] != o).reshape(n, -1).any(1))[0])
            assert np.array_equal(rew[t].view(np.uint32), r.view(np.uint32)) and np.array_equal(done[t], d), (level, rep, t)
            assert np.array_equal(dirs[t], ref.direction), (level, rep, t)
            if mode == 0:
                episodes += int(d.sum())
        if mode == 0:
            steps += T * n
            assert cnt[0] == steps and cnt[1] == episodes and cnt[3] == 0, (cnt, steps, episodes)
        assert r2.error_flag() == 0
        for i in range(0, n, 3):
            g, st = ref.state(i)
            g2, info = r2.state(i, ref.width, ref.height)
            assert np.array_equal(g, g2), (level, rep, i)
            assert (st['agent_x'], st['agent_y'], st['agent_dir'], st['step_count']) == (info[0], info[1], info[2], info[4]), (level, rep, i)
        assert all(np.array_equal(r2.tokens(i)[:8], ref.tokens(i)[:8]) for i in range(n))


@pytest.mark.timeout(300)
@pytest.mark.parametrize('kernel', ['lane', 'pipe'])
def test_rollout_stepping_role_freeze_mode(kernel):
    """ManyEnvs flavour in k_rollout / k_rollout_pipe: finished envs freeze and replay their terminal result (evaluate.py:72-78)."""
    level, n, T = 'GoToLocal', 20, 40
    seeds = np.arange(n, dtype=np.uint64) + 10 ** 9
    ref = _emu(level, n, seeds, mode=1)
    r2 = hostemu.RolloutPool(level_spec(level), n, seeds, mode=1)
    ref.reset()
    rng = np.random.RandomState(3)
    for rep in range(2):                                  # max_steps = 64: everything is frozen during the second rollout
        acts = rng.randint(0, 7, (T, n)).astype(np.int8)
        obs, rew, done, dirs, cnt = r2.rollout(acts, kernel=kernel)
        for t in range(T):
            o, r, d = ref.step(acts[t])
            assert np.array_equal(obs[t], o) and np.array_equal(rew[t].view(np.uint32), r.view(np.uint32)) and np.array_equal(done[t], d), (rep, t)
    assert done[-1].all()


@pytest.mark.timeout(900)
@pytest.mark.parametrize('level,n,T,rounds,min_active', [('GoToLocal', 100, 16, 1 << 20, 0), ('PickupLoc', 70, 16, 2, 16), ('GoToObjS4', 64, 16, 1, 0),
                                                         ('GoToRedBallGrey', 90, 12, 3, 8)])
@pytest.mark.parametrize('kernel', ['lane', 'pipe'])
def test_rollout_fused_cta_on_threads(level, n, T, rounds, min_active, kernel):
    """A whole fused CTA of k_rollout on threads: two stepping warps + the generator warp (rollout_gen_warp driving
    gen_small_round -- the round function of k_gen_small too) behind one __syncthreads.  Nothing
    but the generator warps refills the rings over 14 launches, with small round budgets and the sparse-warp rule switched
    on in some cases (deficits carry over; the must-complete rule keeps every ring above what the next launch can consume)."""
    seeds = np.arange(n, dtype=np.uint64) * 5 + 2024
    ref = _emu(level, n, seeds)
    r2 = hostemu.RolloutPool(level_spec(level), n, seeds, depth=2 * T + 8)       # the smallest ring fused launches accept
    ref.reset()
    rng = np.random.RandomState(23)
    steps = episodes = 0
    for rep in range(14):
        acts = rng.choice(7, size=(T, n), p=[0.13, 0.13, 0.4, 0.12, 0.08, 0.12, 0.02]).astype(np.int8)
        obs, rew, done, dirs, cnt = r2.rollout(acts, fused=True, gen_rounds=rounds, gen_min_active=min_active, kernel=kernel)
        for t in range(T):
            o, r, d = ref.step(acts[t])
            assert np.array_equal(obs[t], o) and np.array_equal(rew[t].view(np.uint32), r.view(np.uint32)) and np.array_equal(done[t], d), (level, rep, t)
            episodes += int(d.sum())
        steps += T * n
        assert cnt[0] == steps and cnt[1] == episodes and cnt[3] == 0, (rep, cnt, steps, episodes)
        assert r2.min_ring_level() >= T, (rep, r2.min_ring_level())
    assert episodes > n
