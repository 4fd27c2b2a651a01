// hostemu.cpp -- TEST-ONLY host build of babyai_b200/csrc/env_logic.cuh.
//
// The product has no CPU path: babyai_b200/ never loads this.  It exists so that
// the `not gpu` test-suite (which runs in a container without a GPU) can execute
// the very same per-environment source the CUDA kernels inline -- generation,
// step, verifier, observation, word staging -- and compare it with the oracle.
// The pool semantics of pool.cu (spare slot, refill, auto-reset / freeze) are
// mirrored sequentially.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#include <stdio.h>
#include "../../babyai_b200/csrc/env_logic.cuh"
#include "../../babyai_b200/csrc/level_params.h"
#include "../../include/babyai_b200.h"

using namespace bb;

struct Slot { std::vector<uint8_t> grid; EnvHot hot; ObjTab obj; InstrRec ins; std::vector<int16_t> tok; };
struct HPool {
    LevelParams lp; int n; int mode;
    std::vector<Slot> live, spare; std::vector<uint8_t> sready, locked_room; std::vector<RngRec> rng;
    std::vector<uint32_t> attempts; std::vector<float> last_reward;
};

static void make_params(const bb_level_spec *s, LevelParams *lp)
{
    const char *e = make_level_params(s, lp);
    if (e) { fprintf(stderr, "hostemu: %s\n", e); abort(); }
    lp->max_tokens = MAXTOK;
}

// env var HOSTEMU_GEN=generic forces the nested-loop generator; default: the flat small-level generator when the
// level qualifies, and in that case BOTH are run and every output byte compared.
static bool records_equal(const LevelParams &lp, Slot &a, Slot &b)
{
    return memcmp(a.grid.data(), b.grid.data(), lp.cells_pad) == 0 && memcmp(&a.hot, &b.hot, sizeof(EnvHot)) == 0 &&
           memcmp(&a.obj, &b.obj, sizeof(ObjTab)) == 0 && memcmp(&a.ins, &b.ins, sizeof(InstrRec)) == 0 &&
           memcmp(a.tok.data(), b.tok.data(), lp.max_tokens * sizeof(int16_t)) == 0;
}
static void gen_spare(HPool *p, int e)
{
    Slot &s = p->spare[e];
    LevelOut o; o.grid = s.grid.data(); o.hot = &s.hot; o.obj = &s.obj; o.ins = &s.ins; o.tok = s.tok.data();
    const char *force = getenv("HOSTEMU_GEN");
    if (p->lp.small && !(force && !strcmp(force, "generic"))) {
        RngRec r0 = p->rng[e];
        RngScalar rng; rng.init(r0.seed, r0.draws);
        SmallLevel L;
        int att = generate_small(p->lp, rng, L);
        memset(s.grid.data(), 0, s.grid.size());
        emit_small_level(p->lp, L, o);
        // cross-check with the nested-loop generator from the same stream position
        Slot ref = s; ref.grid.assign(s.grid.size(), 0);
        LevelOut oref; oref.grid = ref.grid.data(); oref.hot = &ref.hot; oref.obj = &ref.obj; oref.ins = &ref.ins; oref.tok = ref.tok.data();
        GenMemX mem; uint8_t lr = p->locked_room[e]; RngRec r1 = r0;
        int att2 = generate_level(p->lp, oref, &r1, &lr, &mem);
        if (att != att2 || r1.draws != rng.draws || !records_equal(p->lp, s, ref)) {
            fprintf(stderr, "hostemu: generate_small != generate_level (env %d: attempts %d/%d draws %llu/%llu)\n", e, att, att2,
                    (unsigned long long)rng.draws, (unsigned long long)r1.draws);
            abort();
        }
        p->rng[e].draws = rng.draws;
        p->attempts[e] += (uint32_t)att;
    } else {
        GenMemX mem;
        p->attempts[e] += (uint32_t)generate_level(p->lp, o, &p->rng[e], &p->locked_room[e], &mem);
    }
    p->sready[e] = 1;
}

// KIND_UNLOCK pools run the UNTR = true instantiations, as pool.cu does
static int carry_cell(const HPool *p, const EnvHot &h, const GlobalMem &mem)
{
    return p->lp.kind == KIND_UNLOCK ? carry_cell_of<true>(h, mem) : carry_cell_of(h, mem);
}

static void obs_of(HPool *p, int e, uint8_t *out)
{
    Slot &s = p->live[e];
    uint32_t w[OBS_WORDS];
    GlobalMem mem(p->lp, s.grid.data(), &s.obj, &s.ins);
    observe(p->lp, mem, s.hot.x, s.hot.y, s.hot.dirflags & 3, carry_cell(p, s.hot, mem), w);
    memcpy(out, w, OBS_BYTES);
}

extern "C" {

HPool *he_create(const bb_level_spec *spec, int n)
{
    HPool *p = new HPool();
    make_params(spec, &p->lp); p->n = n; p->mode = BB_MODE_AUTORESET;
    p->live.resize(n); p->spare.resize(n); p->sready.assign(n, 0); p->locked_room.assign(n, 0xFF);
    p->rng.resize(n); p->attempts.assign(n, 0); p->last_reward.assign(n, 0.f);
    for (int i = 0; i < n; i++) {
        p->live[i].grid.assign(p->lp.cells_pad, 0); p->spare[i].grid.assign(p->lp.cells_pad, 0);
        p->live[i].tok.assign(MAXTOK, 0); p->spare[i].tok.assign(MAXTOK, 0);
        memset(&p->live[i].hot, 0, sizeof(EnvHot)); memset(&p->live[i].obj, 0, sizeof(ObjTab)); memset(&p->live[i].ins, 0, sizeof(InstrRec));
        p->rng[i].seed = (uint64_t)i; p->rng[i].draws = 0;
    }
    return p;
}
void he_destroy(HPool *p) { delete p; }
void he_set_mode(HPool *p, int mode) { p->mode = mode; }
void he_seed(HPool *p, const uint64_t *seeds)
{
    for (int i = 0; i < p->n; i++) { p->rng[i].seed = seeds[i]; p->rng[i].draws = 0; p->locked_room[i] = 0xFF; p->sready[i] = 0; }
}
void he_reset(HPool *p, uint8_t *obs, int8_t *dir)
{
    for (int e = 0; e < p->n; e++) {
        if (!p->sready[e]) gen_spare(p, e);
        p->live[e] = p->spare[e]; p->sready[e] = 0;
        if (p->mode == BB_MODE_AUTORESET) gen_spare(p, e);
        obs_of(p, e, obs + (size_t)e * OBS_BYTES);
        if (dir) dir[e] = (int8_t)(p->live[e].hot.dirflags & 3);
    }
}
void he_step(HPool *p, const int8_t *actions, uint8_t *obs, float *reward, uint8_t *done, int8_t *dir)
{
    for (int e = 0; e < p->n; e++) {
        Slot &s = p->live[e];
        float rew = 0; bool dn = false;
        if (!(s.hot.dirflags & 4)) {
            GlobalMem mem(p->lp, s.grid.data(), &s.obj, &s.ins);
            StepResult r = p->lp.kind == KIND_UNLOCK ? step_env<true>(s.hot, mem, actions[e]) : step_env(s.hot, mem, actions[e]);
            rew = r.reward; dn = r.done;
            if (dn) {
                if (p->mode == BB_MODE_AUTORESET) { p->live[e] = p->spare[e]; p->sready[e] = 0; gen_spare(p, e); }
                else { s.hot.dirflags |= 4; p->last_reward[e] = rew; }
            }
        } else { rew = p->last_reward[e]; dn = true; }
        obs_of(p, e, obs + (size_t)e * OBS_BYTES);
        reward[e] = rew; done[e] = dn; if (dir) dir[e] = (int8_t)(p->live[e].hot.dirflags & 3);
    }
}
void he_tokens(HPool *p, int e, int16_t *out) { memcpy(out, p->live[e].tok.data(), MAXTOK * sizeof(int16_t)); }
void he_get_state(HPool *p, int e, uint8_t *grid, int32_t *info)
{
    Slot &s = p->live[e];
    const LevelParams &lp = p->lp;
    for (int y = 0; y < lp.H; y++) memcpy(grid + y * lp.W, s.grid.data() + y * lp.rs_g, lp.W);
    for (int y = 0; y < lp.H; y++)          // the column-major copy must mirror the row-major one
        for (int x = 0; x < lp.W; x++)
            if (s.grid[lp.gt_off + x * lp.rs_t + y] != s.grid[y * lp.rs_g + x]) { fprintf(stderr, "hostemu: G/GT mismatch\n"); abort(); }
    {   // and the SWAR observation must equal the cell-by-cell one
        uint32_t w[OBS_WORDS]; uint8_t simple[OBS_BYTES];
        GlobalMem mem(lp, s.grid.data(), &s.obj, &s.ins);
        observe(lp, mem, s.hot.x, s.hot.y, s.hot.dirflags & 3, carry_cell(p, s.hot, mem), w);
        observe_simple(lp, s.grid.data(), s.hot.x, s.hot.y, s.hot.dirflags & 3, carry_cell(p, s.hot, mem), simple);
        if (memcmp(w, simple, OBS_BYTES) != 0) { fprintf(stderr, "hostemu: observe != observe_simple\n"); abort(); }
        uint8_t cols[OBS_BYTES];
        observe_columns(lp, mem, s.hot.x, s.hot.y, s.hot.dirflags & 3, carry_cell(p, s.hot, mem), cols);
        if (memcmp(cols, simple, OBS_BYTES) != 0) { fprintf(stderr, "hostemu: observe_columns != observe_simple\n"); abort(); }
    }
    info[0] = s.hot.x; info[1] = s.hot.y; info[2] = s.hot.dirflags & 3;
    info[3] = s.hot.carry == NO_OBJ ? 0 : (lp.kind == KIND_UNLOCK && (s.hot.carry & CARRY_UNTRACKED)) ? (s.hot.carry & 0x3F) : s.obj.tc[s.hot.carry];
    info[4] = s.hot.step_count; info[5] = s.hot.max_steps;
    info[6] = (int32_t)(p->rng[e].draws & 0x7FFFFFFF); info[7] = (int32_t)p->attempts[e];
}
// single-room levels: observe_room == observe_generic == the literal loops, for every agent pose on env e's live grid
int he_check_room_obs(HPool *p, int e)
{
    Slot &s = p->live[e];
    const LevelParams &lp = p->lp;
    if (!(lp.num_rows == 1 && lp.num_cols == 1)) return -1;
    GlobalMem mem(lp, s.grid.data(), &s.obj, &s.ins);
    int checked = 0;
    for (int y = 1; y < lp.H - 1; y++)
        for (int x = 1; x < lp.W - 1; x++)
            for (int d = 0; d < 4; d++)
                for (int carry = 0; carry < 2; carry++) {
                    const int cc = carry ? (T_KEY | (3 << 3)) : CELL_EMPTY;
                    uint32_t a[OBS_WORDS], b[OBS_WORDS]; uint8_t lit[OBS_BYTES];
                    observe_room(lp, mem, x, y, d, cc, a);
                    observe_generic(lp, mem, x, y, d, cc, b);
                    observe_simple(lp, s.grid.data(), x, y, d, cc, lit);
                    if (memcmp(a, b, OBS_BYTES) != 0 || memcmp(a, lit, OBS_BYTES) != 0) return -2 - checked;
                    checked++;
                }
    return checked;
}
int he_width(HPool *p) { return p->lp.W; }
int he_height(HPool *p) { return p->lp.H; }

// vis_rows against 7-bit see-through rows
void he_vis_rows(const uint32_t *see, uint32_t *vis) { vis_rows(see, vis); }

// staging: 32 lanes x 37 words -> 4704-byte tile, emulating the shuffle
void he_stage(const uint32_t *w /* [32][37] */, uint8_t *tile /* 4704 + slack */)
{
    uint32_t t[32 * OBS_BYTES / 4 + 2];
    memset(t, 0xEE, sizeof t);
    for (int lane = 0; lane < 32; lane++) {
        uint32_t next_w0 = lane < 31 ? w[(lane + 1) * OBS_WORDS] : w[lane * OBS_WORDS];   // shfl_down keeps own value at the edge
        stage_obs_words(t, w + lane * OBS_WORDS, lane, next_w0);
    }
    memcpy(tile, t, 32 * OBS_BYTES);
}

// staging of 28 records of 21 bytes (4 envs x 7 view columns of one warp) -> 588-byte tile
void he_stage21(const uint32_t *w /* [28][6] */, uint8_t *tile)
{
    uint32_t t[147 + 2];
    memset(t, 0xEE, sizeof t);
    for (int q = 0; q < 28; q++) {
        uint32_t next_w0 = q < 27 ? w[(q + 1) * 6] : 0u;
        stage_record_words<21, 6>(t, w + q * 6, q, next_w0);
    }
    memcpy(tile, t, 588);
}

}  // extern "C"
extern "C" int he_verifier_modes(HPool *p) { return p->lp.strict_mask | (p->lp.done_actions << 8); }
