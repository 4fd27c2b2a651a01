// simt_rollout.cpp -- TEST-ONLY: runs the stepping roles of the rollout kernels (babyai_b200/csrc/rollout_lane.cuh and
// rollout_cta.cuh: the very functions k_rollout / k_rollout_cta call) and the fused generator warp (gen_round.cuh) on the
// host with ONE OS THREAD PER LANE.  The warp primitives the functions use
// (__syncwarp, __shfl_sync, __shfl_xor_sync, __shfl_down_sync) are rendezvous on a per-warp barrier; every lane of a warp
// reaches them in the same order (none sits inside divergent code), which is exactly what the hardware requires too --
// a lane that skipped one would dead-lock here instead of silently reading garbage.  Global memory is host memory laid
// out as pool.cu lays it out (struct of arrays + the per-env ring of pre-generated levels); shared memory is a per-warp
// buffer.  The GPU-less suite steps whole rollouts through this and compares them with the per-step path.
//
// The product never loads this file.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <barrier>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v = { x, y, z, w }; return v; }

// ---- the warp primitives of simt.cuh, emulated ---------------------------------------------------------------
struct WarpCtx {
    std::barrier<> bar{32};
    std::barrier<> grp0{8}, grp1{8}, grp2{8}, grp3{8};       // __syncwarp(mask) of k_step8's 8-lane groups
    uint32_t xchg[32];
};
static thread_local WarpCtx *tl_warp = nullptr;
static thread_local int tl_lane = 0;
static thread_local std::barrier<> *tl_cta = nullptr;       // __syncthreads of a fused launch (4 stepping warps + the generator warp)

// named barriers of a CTA (barrier.arrive / barrier.sync with a thread count): a phase completes when `n` threads have
// arrived or synced; arrive returns at once, sync waits for the phase to complete
struct NamedBar { std::mutex m; std::condition_variable cv; int count = 0; unsigned gen = 0; };
static thread_local NamedBar *tl_named = nullptr;           // 8 per CTA
static inline void emu_bar_arrive(int id, int n)
{
    NamedBar &b = tl_named[id];
    std::unique_lock<std::mutex> lk(b.m);
    if (++b.count == n) { b.count = 0; b.gen++; b.cv.notify_all(); }
}
static inline void emu_bar_sync(int id, int n)
{
    NamedBar &b = tl_named[id];
    std::unique_lock<std::mutex> lk(b.m);
    const unsigned g = b.gen;
    if (++b.count == n) { b.count = 0; b.gen++; b.cv.notify_all(); }
    else b.cv.wait(lk, [&] { return b.gen != g; });
}

static thread_local int *tl_cta_or = nullptr;               // two alternating accumulators of __syncthreads_or
static thread_local int tl_cta_phase = 0;
static inline void emu_syncwarp() { tl_warp->bar.arrive_and_wait(); }
static inline void emu_syncwarp_mask(unsigned mask)         // the 8 lanes of one group (mask = 0xFF << 8 g)
{
    const int g = __builtin_ctz(mask) >> 3;
    (g == 0 ? tl_warp->grp0 : g == 1 ? tl_warp->grp1 : g == 2 ? tl_warp->grp2 : tl_warp->grp3).arrive_and_wait();
}
static inline int emu_syncthreads_or(int pred)
{
    int *acc = tl_cta_or + (tl_cta_phase & 1);
    if (pred) __atomic_store_n(acc, 1, __ATOMIC_RELAXED);
    tl_cta->arrive_and_wait();
    const int r = __atomic_load_n(acc, __ATOMIC_RELAXED);
    tl_cta->arrive_and_wait();                           // everybody has read it ...
    if (tl_lane == 0) __atomic_store_n(tl_cta_or + ((tl_cta_phase + 1) & 1), 0, __ATOMIC_RELAXED);   // ... the OTHER one is clear for the next use
    tl_cta_phase++;
    return r;
}
static inline uint32_t emu_exchange(uint32_t v, int src)
{
    tl_warp->xchg[tl_lane] = v;
    tl_warp->bar.arrive_and_wait();
    const uint32_t r = tl_warp->xchg[src & 31];
    tl_warp->bar.arrive_and_wait();                      // nobody overwrites xchg before everybody has read it
    return r;
}
static inline uint32_t emu_ballot(bool pred)
{
    tl_warp->xchg[tl_lane] = pred ? 1u : 0u;
    tl_warp->bar.arrive_and_wait();
    uint32_t m = 0;
    for (int l = 0; l < 32; l++) m |= tl_warp->xchg[l] << l;
    tl_warp->bar.arrive_and_wait();
    return m;
}
template <class T> static inline T emu_shfl(T v, int src) { return (T)emu_exchange((uint32_t)v, src); }
template <class T> static inline T emu_shfl_down(T v, int d) { return (T)emu_exchange((uint32_t)v, tl_lane + d < 32 ? tl_lane + d : tl_lane); }
template <class T> static inline T emu_shfl_xor(T v, int m) { return (T)emu_exchange((uint32_t)v, tl_lane ^ m); }

#define BB_DEV inline
#define BB_SYNCWARP() emu_syncwarp()
#define BB_SYNCTHREADS() tl_cta->arrive_and_wait()
#define BB_SYNCWARP_MASK(m) emu_syncwarp_mask(m)
#define BB_CP_ASYNC16(dst, src) memcpy((dst), (src), 16)      /* cp.async: a plain copy here; the wait is a no-op */
#define BB_CP_ASYNC_WAIT_ALL() ((void)0)
#define BB_ANY(x) (emu_ballot(x) != 0u)
#define BB_BALLOT(x) emu_ballot(x)
#define BB_POPC(x) __builtin_popcount(x)
#define BB_SHFL(v, src) emu_shfl((v), (src))
#define BB_SHFL_XOR(v, m) emu_shfl_xor((v), (m))
#define BB_SHFL_DOWN(v, d) emu_shfl_down((v), (d))
#define BB_LDCG(p) (*(p))
#define BB_ATOMIC_ADD(p, v) __atomic_fetch_add((p), (v), __ATOMIC_RELAXED)
#define BB_PREFETCH_L2(p) ((void)(p))
#define BB_LD_S8(p) ((int)*(p))
#define BB_SYNCTHREADS_OR(x) emu_syncthreads_or(x)
#define BB_ROLE_SYNC(n) tl_cta->arrive_and_wait()             /* the named barrier of the kernel's warp roles */
// the bulk (async proxy) tile store: a plain copy here; the fence / wait are no-ops
#define BB_FENCE_ASYNC_SMEM() ((void)0)
#define BB_BULK_STORE(gdst, ssrc, bytes) memcpy((gdst), (ssrc), (bytes))
#define BB_BULK_WAIT_READ() ((void)0)
#define BB_BULK_WAIT_READ_N(n) ((void)0)
#define BB_BAR_SYNC(id, n) emu_bar_sync((id), (n))
#define BB_BAR_ARRIVE(id, n) emu_bar_arrive((id), (n))

#include "../../babyai_b200/csrc/simt.cuh"
#include "../../babyai_b200/csrc/gen_round.cuh"
#include "../../babyai_b200/csrc/rollout_lane.cuh"
#include "../../babyai_b200/csrc/rollout_cta.cuh"
#include "../../babyai_b200/csrc/step8.cuh"
#include "../../babyai_b200/csrc/level_params.h"

using namespace bb;

// the accessor k_rollout uses on shared memory (pool.cu SmemOnlyMem), restated for the host
struct HostSmemMem {
    static constexpr bool untracked = false;
    const LevelParams &lp; uint8_t *g, *o, *i;
    HostSmemMem(const LevelParams &lp_, uint8_t *g_, uint8_t *o_, uint8_t *i_) : lp(lp_), g(g_), o(o_), i(i_) {}
    int cell(int x, int y) const { return g[y * lp.rs_g + x]; }
    void set_cell(int x, int y, int v) { bb::set_cell(lp, g, x, y, v); }
    uint32_t word_at(int off) const { uint32_t v; memcpy(&v, g + off, 4); return v; }
    int ox(int k) const { return o[k]; }
    int oy(int k) const { return o[MAXOBJ + k]; }
    int otc(int k) const { return o[2 * MAXOBJ + k]; }
    uint32_t oxw(int i) const { uint32_t v; memcpy(&v, o + 4 * i, 4); return v; }
    uint32_t oyw(int i) const { uint32_t v; memcpy(&v, o + MAXOBJ + 4 * i, 4); return v; }
    void set_oxy(int k, int x, int y) { o[k] = (uint8_t)x; o[MAXOBJ + k] = (uint8_t)y; }
    uint32_t desc_mask(int d) const { uint32_t v; memcpy(&v, i + 4 * d, 4); return v; }
    int leaf_kind(int l) const { return i[32 + l]; }
    int leaf_pre(int l) const { return i[36 + l]; }
    void set_leaf_pre(int l, int v) { i[36 + l] = (uint8_t)v; }
    int root_kind() const { return i[40]; }
    int side_and() const { return i[41]; }
    void set_side_and(int v) { i[41] = (uint8_t)v; }
    int flags() const { return i[42]; }
    void set_flags(int v) { i[42] = (uint8_t)v; }
    int start_carry() const { return i[43]; }
};

template <int K>
struct HostSmemRoomKindMem : HostSmemMem {                // the single-instruction-kind single-room instantiations (env_logic.cuh mem_spec), as pool.cu picks them
    static constexpr int spec_room_kinds = K;
    HostSmemRoomKindMem(const LevelParams &lp_, uint8_t *g_, uint8_t *o_, uint8_t *i_) : HostSmemMem(lp_, g_, o_, i_) {}
};

struct HostPoolPtrs {                                     // the members of pool.cu's PoolPtrs the stepping role touches
    uint8_t *grid; EnvHot *hot; ObjTab *obj; InstrRec *ins; int16_t *tok;
    uint8_t *rgrid; EnvHot *rhot; ObjTab *robj; InstrRec *rins; int16_t *rtok;
    uint32_t *head, *tail, *tail_pub;
    RngRec *rng; uint32_t *attempts;
    float *last_reward;
    unsigned long long *warp_counters;
    int *err_flag;
    int32_t depth, n;
};

struct RPool {
    LevelParams lp; int n, D, mode;
    std::vector<uint8_t> grid, rgrid, locked_room;
    std::vector<EnvHot> hot, rhot; std::vector<ObjTab> obj, robj; std::vector<InstrRec> ins, rins;
    std::vector<int16_t> tok, rtok;
    std::vector<uint32_t> head, tail, tail_pub, attempts;
    std::vector<RngRec> rng; std::vector<float> last_reward;
    std::vector<unsigned long long> counters;
    int err;
    HostPoolPtrs P;
};

static LevelOut slot_of(RPool *p, int env, int slot) { return r2_ring_slot(p->lp, p->P, env, slot); }

static void refill(RPool *p)                              // k_gen's job: top every ring up to D levels
{
    for (int e = 0; e < p->n; e++)
        while ((int)(p->tail[e] - p->head[e]) < p->D) {
            GenMemX mem;
            generate_level(p->lp, slot_of(p, e, (int)(p->tail[e] % (uint32_t)p->D)), &p->rng[e], &p->locked_room[e], &mem);
            p->tail[e]++;
        }
    p->tail_pub = p->tail;
    p->P.tail_pub = p->tail_pub.data();
}

extern "C" {

RPool *r2_create(const bb_level_spec *spec, int n, int depth, const uint64_t *seeds, int mode)
{
    RPool *p = new RPool();
    const char *err = make_level_params(spec, &p->lp);
    if (err) { fprintf(stderr, "simt_rollout: %s\n", err); abort(); }
    const LevelParams &lp = p->lp;
    p->n = n; p->D = depth; p->mode = mode;
    const size_t N = (size_t)n, DN = (size_t)depth * n;
    p->grid.assign(N * lp.cells_pad, 0); p->rgrid.assign(DN * lp.cells_pad, 0);
    p->hot.resize(N); p->rhot.resize(DN); p->obj.resize(N); p->robj.resize(DN); p->ins.resize(N); p->rins.resize(DN);
    memset(p->hot.data(), 0, N * sizeof(EnvHot)); memset(p->obj.data(), 0, N * sizeof(ObjTab)); memset(p->ins.data(), 0, N * sizeof(InstrRec));
    p->tok.assign(N * lp.max_tokens, 0); p->rtok.assign(DN * lp.max_tokens, 0);
    p->head.assign(N, 0); p->tail.assign(N, 0); p->tail_pub.assign(N, 0);
    p->rng.resize(N); p->last_reward.assign(N, 0.f); p->locked_room.assign(N, 0xFF); p->attempts.assign(N, 0);
    p->counters.assign(4 * (N / 4 + 8), 0); p->err = 0;
    for (int e = 0; e < n; e++) { p->rng[e].seed = seeds[e]; p->rng[e].draws = 0; }
    HostPoolPtrs &P = p->P;
    P.grid = p->grid.data(); P.hot = p->hot.data(); P.obj = p->obj.data(); P.ins = p->ins.data(); P.tok = p->tok.data();
    P.rgrid = p->rgrid.data(); P.rhot = p->rhot.data(); P.robj = p->robj.data(); P.rins = p->rins.data(); P.rtok = p->rtok.data();
    P.head = p->head.data(); P.tail = p->tail.data(); P.tail_pub = p->tail_pub.data();
    P.last_reward = p->last_reward.data(); P.warp_counters = p->counters.data();
    P.rng = p->rng.data(); P.attempts = p->attempts.data(); P.err_flag = &p->err;
    P.depth = depth; P.n = n;
    // reset: generate, then take the first level of every ring as the live state (what bb_pool_reset does)
    refill(p);
    for (int e = 0; e < n; e++) {
        const LevelOut o = slot_of(p, e, 0);
        memcpy(P.grid + (size_t)e * lp.cells_pad, o.grid, lp.cells_pad);
        P.hot[e] = *o.hot; P.obj[e] = *o.obj; P.ins[e] = *o.ins;
        memcpy(P.tok + (size_t)e * lp.max_tokens, o.tok, lp.max_tokens * sizeof(int16_t));
        p->head[e] = 1;
    }
    refill(p);
    return p;
}
void r2_destroy(RPool *p) { delete p; }

// one bb_pool_rollout worth of k_rollout (non-fused launch): every warp of the grid, 32 threads each
void r2_rollout(RPool *p, const int8_t *actions, int T, uint8_t *obs, float *reward, uint8_t *done, int8_t *dirs, int64_t *counters4)
{
    const LevelParams &lp = p->lp;
    const int warp_words = rl_warp_words(lp);
    const int nwarps = (p->n + 31) / 32;
    for (int wg = 0; wg < nwarps; wg++) {
        WarpCtx ctx;
        std::vector<uint32_t> smem((size_t)warp_words + 8, 0xDEADBEEFu);
        uint32_t *base = smem.data();
        while (((uintptr_t)base) & 15) base++;                // the kernel's tile is 16-byte aligned
        std::vector<std::thread> th;
        for (int lane = 0; lane < 32; lane++)
            th.emplace_back([&, lane]() {
                tl_warp = &ctx; tl_lane = lane;
                if (lp.kind == KIND_UNLOCK) rollout_lane_step_warp<HostPoolPtrs, HostSmemMem, 1, true>(lp, p->P, actions, obs, reward, done, dirs, p->n, T, p->mode, 0, false, base, lane, wg, nullptr);
                else if (level_spec_room_kinds(lp) == (1 << I_GOTO)) rollout_lane_step_warp<HostPoolPtrs, HostSmemRoomKindMem<1 << I_GOTO>, 1, false>(lp, p->P, actions, obs, reward, done, dirs, p->n, T, p->mode, 0, false, base, lane, wg, nullptr);
                else if (level_spec_room_kinds(lp) == (1 << I_PICKUP)) rollout_lane_step_warp<HostPoolPtrs, HostSmemRoomKindMem<1 << I_PICKUP>, 1, false>(lp, p->P, actions, obs, reward, done, dirs, p->n, T, p->mode, 0, false, base, lane, wg, nullptr);
                else rollout_lane_step_warp<HostPoolPtrs, HostSmemMem, 1, false>(lp, p->P, actions, obs, reward, done, dirs, p->n, T, p->mode, 0, false, base, lane, wg, nullptr);
            });
        for (auto &t : th) t.join();
    }
    refill(p);                                                // the refill pass between launches
    for (int k = 0; k < 4; k++) counters4[k] = 0;
    for (size_t w = 0; w < p->counters.size() / 4; w++) for (int k = 0; k < 4; k++) counters4[k] += (int64_t)p->counters[4 * w + k];
}

// a FUSED launch of k_rollout: per CTA two stepping warps + the generator warp (rollout_gen_warp, gen_small_round), 96
// threads, one CTA after the other; nothing refills the rings but the generator warps
void r2_rollout_fused(RPool *p, const int8_t *actions, int T, int gen_rounds, int gen_min_active, uint8_t *obs, float *reward,
                      uint8_t *done, int8_t *dirs, int64_t *counters4)
{
    const LevelParams &lp = p->lp;
    if (!lp.small) { fprintf(stderr, "simt_rollout: fused launches are for small single-room levels\n"); abort(); }
    const int warp_words = rl_warp_words(lp);
    const int SW = 2, cta_envs = SW * 32;
    const int nctas = (p->n + cta_envs - 1) / cta_envs;
    for (int cta = 0; cta < nctas; cta++) {
        std::vector<WarpCtx> ctx(SW + 1);
        std::barrier<> cta_bar(32 * (SW + 1));
        std::vector<uint32_t> smem((size_t)SW * warp_words + RG_AREA_WORDS + 8, 0xDEADBEEFu);
        uint32_t *smr = smem.data();
        while (((uintptr_t)smr) & 15) smr++;
        uint32_t *g_area = smr + SW * warp_words;
        volatile int *s_done = reinterpret_cast<volatile int *>(g_area + RG_AREA_WORDS - 4);
        std::vector<std::thread> th;
        for (int tid = 0; tid < 32 * (SW + 1); tid++)
            th.emplace_back([&, tid]() {
                const int lane = tid & 31, warp = tid >> 5;
                tl_warp = &ctx[warp]; tl_lane = lane; tl_cta = &cta_bar;
                if (warp == SW) rollout_gen_warp(lp, p->P, g_area, s_done, p->n, T, cta * cta_envs, gen_rounds, gen_min_active, lane, SW);
                else if (level_spec_room_kinds(lp) == (1 << I_GOTO)) rollout_lane_step_warp<HostPoolPtrs, HostSmemRoomKindMem<1 << I_GOTO>, 1, false>(lp, p->P, actions, obs, reward, done, dirs, p->n, T, p->mode, 0, true,
                                                                                 smr + warp * warp_words, lane, cta * SW + warp, s_done);
                else if (level_spec_room_kinds(lp) == (1 << I_PICKUP)) rollout_lane_step_warp<HostPoolPtrs, HostSmemRoomKindMem<1 << I_PICKUP>, 1, false>(lp, p->P, actions, obs, reward, done, dirs, p->n, T, p->mode, 0, true,
                                                                                 smr + warp * warp_words, lane, cta * SW + warp, s_done);
                else rollout_lane_step_warp<HostPoolPtrs, HostSmemMem, 1, false>(lp, p->P, actions, obs, reward, done, dirs, p->n, T, p->mode, 0, true,
                                                                                 smr + warp * warp_words, lane, cta * SW + warp, s_done);
            });
        for (auto &t : th) t.join();
    }
    for (int k = 0; k < 4; k++) counters4[k] = 0;
    for (size_t w = 0; w < p->counters.size() / 4; w++) for (int k = 0; k < 4; k++) counters4[k] += (int64_t)p->counters[4 * w + k];
}

// one bb_pool_rollout worth of k_rollout_cta: every CTA of the grid, 128 threads each (4 warps: shuffles within a warp,
// __syncthreads / __syncthreads_or across the CTA)
void r2_rollout_cta(RPool *p, const int8_t *actions, int T, uint8_t *obs, float *reward, uint8_t *done, int8_t *dirs, int64_t *counters4)
{
    const LevelParams &lp = p->lp;
    const int words = rc2_cta_words(lp);
    const int nctas = (p->n + RC_ENVS - 1) / RC_ENVS;
    for (int cta = 0; cta < nctas; cta++) {
        std::vector<WarpCtx> ctx(RC_THREADS / 32);
        std::barrier<> cta_bar(RC_THREADS);
        int cta_or[2] = { 0, 0 };
        std::vector<NamedBar> named(8);
        std::vector<uint32_t> smem((size_t)words + 8, 0xDEADBEEFu);
        uint32_t *base = smem.data();
        while (((uintptr_t)base) & 15) base++;
        std::vector<std::thread> th;
        for (int tid = 0; tid < RC_THREADS; tid++)
            th.emplace_back([&, tid]() {
                tl_warp = &ctx[tid >> 5]; tl_lane = tid & 31; tl_cta = &cta_bar; tl_cta_or = cta_or; tl_cta_phase = 0; tl_named = named.data();
                if (lp.kind == KIND_UNLOCK) rollout_cta_role<HostPoolPtrs, true>(lp, p->P, actions, obs, reward, done, dirs, p->n, T, p->mode, base, tid, cta);
                else rollout_cta_role<HostPoolPtrs, false>(lp, p->P, actions, obs, reward, done, dirs, p->n, T, p->mode, base, tid, cta);
            });
        for (auto &t : th) t.join();
    }
    refill(p);                                                // the generation pass between launches
    for (int k = 0; k < 4; k++) counters4[k] = 0;
    for (size_t w = 0; w < p->counters.size() / 4; w++) for (int k = 0; k < 4; k++) counters4[k] += (int64_t)p->counters[4 * w + k];
}

// one bb_pool_step worth of k_step8: every CTA of the grid, 128 threads each (4 warps of four 8-lane groups; shuffles, ballots and
// the sub-warp __syncwarp(mask) as rendezvous), then the generation pass the pool's scheduler would have run by now
void r2_step8(RPool *p, const int8_t *actions, uint8_t *obs, float *reward, uint8_t *done, int8_t *dirs, int force_reset, int64_t *counters4)
{
    const LevelParams &lp = p->lp;
    const size_t sm8 = (size_t)S8_WARPS * 4 * (lp.cells_pad + S8_REC_FIXED) + (size_t)S8_WARPS * (S8_TILE_WORDS + 1) * 4;
    const int nctas = (p->n + 4 * S8_WARPS - 1) / (4 * S8_WARPS);
    for (int cta = 0; cta < nctas; cta++) {
        std::vector<WarpCtx> ctx(S8_WARPS);
        std::vector<uint8_t> smem(sm8 + 32, 0xEE);
        uint8_t *base = smem.data();
        while (((uintptr_t)base) & 15) base++;
        std::vector<std::thread> th;
        for (int tid = 0; tid < S8_THREADS; tid++)
            th.emplace_back([&, tid]() {
                tl_warp = &ctx[tid >> 5]; tl_lane = tid & 31;
                if (lp.kind == KIND_UNLOCK) step8_role<HostPoolPtrs, 1, true>(lp, p->P, actions, obs, reward, done, dirs, p->n, p->mode, force_reset, base, tid & 31, tid >> 5, (unsigned)cta);
                else step8_role<HostPoolPtrs, 1, false>(lp, p->P, actions, obs, reward, done, dirs, p->n, p->mode, force_reset, base, tid & 31, tid >> 5, (unsigned)cta);
            });
        for (auto &t : th) t.join();
    }
    refill(p);
    for (int k = 0; k < 4; k++) counters4[k] = 0;
    for (size_t w = 0; w < p->counters.size() / 4; w++) for (int k = 0; k < 4; k++) counters4[k] += (int64_t)p->counters[4 * w + k];
}

// the live state of env e (both grid orientations must agree after a launch)
void r2_state(RPool *p, int e, uint8_t *grid, int32_t *info)
{
    const LevelParams &lp = p->lp;
    const uint8_t *g = p->grid.data() + (size_t)e * lp.cells_pad;
    for (int y = 0; y < lp.H; y++)
        for (int x = 0; x < lp.W; x++) {
            grid[y * lp.W + x] = g[y * lp.rs_g + x];
            if (g[lp.gt_off + x * lp.rs_t + y] != g[y * lp.rs_g + x]) { fprintf(stderr, "simt_rollout: G/GT mismatch env %d (%d,%d)\n", e, x, y); abort(); }
        }
    const EnvHot &h = p->hot[e];
    info[0] = h.x; info[1] = h.y; info[2] = h.dirflags & 3; info[3] = h.carry == NO_OBJ ? 0 : (lp.kind == KIND_UNLOCK && (h.carry & CARRY_UNTRACKED)) ? (h.carry & 0x3F) : p->obj[e].tc[h.carry];   // the carried object's cell byte, as hostemu.cpp reports it
    info[4] = h.step_count; info[5] = h.max_steps;
}

int r2_min_ring_level(RPool *p) { int m = 1 << 30; for (int e = 0; e < p->n; e++) { const int have = (int)(p->tail[e] - p->head[e]); if (have < m) m = have; } return m; }

int r2_error_flag(RPool *p) { return p->err; }
void r2_tokens(RPool *p, int e, int16_t *out) { memcpy(out, p->tok.data() + (size_t)e * p->lp.max_tokens, p->lp.max_tokens * sizeof(int16_t)); }
int r2_max_tokens(RPool *p) { return p->lp.max_tokens; }

}  // extern "C"
