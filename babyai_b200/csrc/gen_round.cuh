// gen_round.cuh -- one round of the small-level generator for the 32 lanes of a warp (k_gen_small and the generator warps
// of k_rollout in pool.cu).  The warp vote is a macro so that tests/hostemu can compile this very function
// for the host with one OS thread per lane (tests/hostemu/simt_rollout.cpp).
#pragma once
#include "simt.cuh"

namespace bb {

// Generator warp of a fused launch: draw rings of 8 Philox blocks per lane + the work list, in the CTA's shared memory
typedef DrawRingT<8> RolloutRing;
constexpr int RG_RING_WORDS = 32 * RolloutRing::RING_WORDS;            // 1024 words
constexpr int RG_AREA_WORDS = RG_RING_WORDS + 64 /*tails*/ + 32 /*deficits, u16*/ + 16 /*list, u8*/ + 4 /*flag*/;

// One round of the small-level generator for the 32 lanes of a warp (see the comment above k_gen_small): every
// working lane (`active`) makes one attempt at the next level of its env; on success the level is written to ring
// slot tl % D and tl / left advance.  Called with all 32 lanes.
template <class DS, bool PUBLISH, class PP>
BB_DEV void gen_small_round(const LevelParams &lp, const PP &P, DS &ds, const bool active,
                                                const int env, uint32_t &tl, int &left, const uint32_t D)
{
    // converged top-up: when a lane that is about to draw has fewer than RING_LOW draws ready, EVERY working lane
    // generates the blocks its ring has room for
#define BB_TOPUP(cond)                                                                                   \
    if (BB_ANY((cond) && ds.avail() < RING_LOW)) {                                             \
        for (;;) {                                                                                       \
            const bool rm = active && ds.room();                                                         \
            if (!BB_ANY(rm)) break;                                                            \
            if (rm) ds.gen_block();                                                                      \
        }                                                                                                \
    }
    // ---- one attempt per working lane ---------------------------------------------------------------
    SmallAttempt a;
    a.stage = ST_IDLE; a.occ = 0; a.fill = 0; a.k = 0; a.tries = 0; a.cur_tc = 0; a.agent_placed = false;
    a.L.poss = 0; a.L.tcs = 0; a.L.nobj = 0; a.L.ax = a.L.ay = a.L.adir = 0;
    a.L.leaf_kind = 0; a.L.d_type = 0; a.L.d_color = 0; a.L.d_loc = 0; a.L.d_mask = 0;
    if (active) small_attempt_begin(lp, a, ds);
    for (;;) {                                                  // placements: agent and objects, one try per trip
        const bool placing = a.stage == ST_OBJ || a.stage == ST_AGENT;
        if (!BB_ANY(placing)) break;
        BB_TOPUP(placing)
        if (placing) small_place_try(lp, a, ds);
    }
    bool ok = a.stage == ST_PLACED;
    if (small_needs_check(lp)) {                                // check_objs_reachable
        if (ok) small_flood_begin(a);
        for (;;) {
            const bool changed = ok && small_flood_sweep(a);
            if (!BB_ANY(changed)) break;
        }
        ok = ok && small_flood_ok(lp, a);
    }
    a.tries = 0;
    if (lp.kind == KIND_LEVELGEN) {                             // rand_obj: rejection sampling of a descriptor
        bool trying = ok;
        for (;;) {
            if (!BB_ANY(trying)) break;
            BB_TOPUP(trying)
            if (trying && small_desc_try(lp, a, ds)) trying = false;
        }
        ok = ok && a.stage != ST_FAIL;
    } else {
        BB_TOPUP(ok)
        if (ok) small_pick(lp, a, ds);
    }
    // ---- write the level; the env's records are consistent after every round -------------------------
    if (ok) {
        emit_small_level(lp, a.L, r2_ring_slot(lp, P, env, (int)(tl % D)));
        tl++; left--;
        P.tail[env] = tl;
        if (PUBLISH) P.tail_pub[env] = tl;        // generator warp inside k_rollout: nobody reads tail_pub during the launch
    }
    if (active) { P.rng[env].draws = ds.draws; P.attempts[env] += 1u; }
#undef BB_TOPUP
}

// The generator warp of a fused k_rollout launch, for the CTA's envs [cta_env0, cta_env0 + 64) (`step_warps` stepping
// warps count themselves into *s_done when they are finished).  g_area: RG_AREA_WORDS of the CTA's shared memory; s_done: its last word (the stepping warps
// count themselves in there when they are finished).
template <class PP>
BB_DEV void rollout_gen_warp(const LevelParams &lp, const PP &P, uint32_t *g_area, volatile int *s_done, const int n, const int T,
                             const int cta_env0, const int gen_rounds, const int gen_min_active, const int lane, const int step_warps)
{
    const uint32_t D = (uint32_t)P.depth;
    uint32_t *ring = g_area, *s_tl = g_area + RG_RING_WORDS;
    uint16_t *s_def = reinterpret_cast<uint16_t *>(s_tl + 64);
    uint8_t *list = reinterpret_cast<uint8_t *>(s_tl + 64 + 32);
    int cta_nv = n - cta_env0; cta_nv = cta_nv > 64 ? 64 : (cta_nv < 0 ? 0 : cta_nv);
    int cnt = 0;
    bool urgent = false;
    if (lane == 0) *s_done = 0;
#pragma unroll
    for (int h2 = 0; h2 < 2; h2++) {
        const int i = 32 * h2 + lane;
        bool need = false;
        if (i < cta_nv) {
            const uint32_t hd = P.head[cta_env0 + i], tl0 = P.tail[cta_env0 + i];
            const int have = (int)(tl0 - hd);
            s_tl[i] = tl0; s_def[i] = (uint16_t)((int)D - have);
            need = have < (int)D;
            urgent = urgent || have < 2 * T;
        }
        const uint32_t m = BB_BALLOT(need);
        if (need) list[cnt + BB_POPC(m & ((1u << lane) - 1u))] = (uint8_t)i;
        cnt += BB_POPC(m);
    }
    const bool must_complete = BB_ANY(urgent);
    BB_ROLE_SYNC(32 * (step_warps + 1));               // the stepping warps have read head / tail: generation may start
    RolloutRing ds;
    ds.init(ring + lane, 32, 0, 0);
    int env_g = -1, left = 0, next = 0, rounds = 0;
    uint32_t tl = 0;
    for (;;) {
        const bool idle = left == 0;
        const uint32_t midle = BB_BALLOT(idle);
        if (midle && next < cnt) {
            const int idx = next + BB_POPC(midle & ((1u << lane) - 1u));
            if (idle && idx < cnt) {
                const int i = list[idx];
                env_g = cta_env0 + i; tl = s_tl[i]; left = (int)s_def[i];
                const RngRec r = P.rng[env_g];
                ds.init(ring + lane, 32, r.seed, r.draws);
            }
            next += BB_POPC(midle);
        }
        const bool active = left > 0;
        const uint32_t mact = BB_BALLOT(active);
        if (!mact) break;
        if (!must_complete) {
            int dn = 0;
            if (lane == 0) dn = *s_done;
            dn = BB_SHFL(dn, 0);
            if (rounds >= gen_rounds || dn >= step_warps) break;
            if (rounds >= 1 && BB_POPC(mact) < gen_min_active) break;
        }
        rounds++;
        gen_small_round<RolloutRing, true>(lp, P, ds, active, env_g, tl, left, D);
    }
}

}  // namespace bb
