// step8.cuh -- the role function of k_step8 (pool.cu): the one-launch-per-step kernel for multi-room levels, EIGHT LANES PER
// ENVIRONMENT.  Like rollout_lane.cuh / rollout_cta.cuh its warp primitives are the macros of simt.cuh, so that tests/hostemu
// compiles this very function for the host with one OS thread per lane (tests/hostemu/simt_rollout.cpp: r2_step8).
#pragma once
#include "simt.cuh"

namespace bb {

// ---- column-parallel variant: EIGHT LANES PER ENVIRONMENT (default) ------------------------------
// The lane-per-env kernels above are latency-bound: 65 536 envs are only 2 048 warps, 11-14 per SM, and
// each lane carries a ~1 500-2 000 instruction dependent program (r01d/r01h profiles: 0.9 IPC per SM, 9-12
// cycles per issued instruction).  Here a group of 8 lanes serves one env (4 envs per warp, 16 384 warps at
// 65 536 envs):
//   lane 0 of the group applies the action and runs the verifier (step_env) and broadcasts the new pose;
//   lanes 0..6 each fetch ONE view column (3 aligned words) and compute its see-through bits;
//   seven warp ballots turn the column bits of all four envs into row masks (the 8x8 bit transposes of the
//   scalar path for free); every lane runs the 7-row visibility propagation; lanes 0..6 encode their
//   column's 21 output bytes and stage them as aligned words (same funnel-shift scheme, 21-byte records);
//   the warp's 588 observation bytes leave as coalesced 32-bit stores.
//   A finished env's next level is copied from the ring by the 8 lanes together.
constexpr int S8_THREADS = 128;
constexpr int S8_WARPS = S8_THREADS / 32;
constexpr int S8_TILE_WORDS = 4 * OBS_BYTES / 4;               // 147 words: 4 envs per warp
constexpr int S8_REC_FIXED = (int)(sizeof(ObjTab) + sizeof(InstrRec));   // 144 bytes after the grid

// Environment memory of the 8-lane kernel: the env's grid, object table and instruction record staged in
// shared memory (plain byte/word accesses, ~30 cycles), every write mirrored to the state in global memory.
struct StagedMem {
    const LevelParams &lp; uint8_t *sg; ObjTab *so; InstrRec *si;      // shared-memory copies
    uint8_t *grid; ObjTab *ot; InstrRec *ins;                          // global state
    BB_DEV StagedMem(const LevelParams &lp_, uint8_t *sg_, ObjTab *so_, InstrRec *si_, uint8_t *g, ObjTab *o, InstrRec *i)
        : lp(lp_), sg(sg_), so(so_), si(si_), grid(g), ot(o), ins(i) {}
    BB_DEV int cell(int x, int y) const { return sg[y * lp.rs_g + x]; }
    BB_DEV void set_cell(int x, int y, int v) { bb::set_cell(lp, sg, x, y, v); bb::set_cell(lp, grid, x, y, v); }
    BB_DEV uint32_t word_at(int off) const { return *reinterpret_cast<const uint32_t *>(sg + off); }
    BB_DEV int ox(int k) const { return so->x[k]; }
    BB_DEV int oy(int k) const { return so->y[k]; }
    BB_DEV int otc(int k) const { return so->tc[k]; }
    BB_DEV uint32_t oxw(int i) const { return reinterpret_cast<const uint32_t *>(so->x)[i]; }
    BB_DEV uint32_t oyw(int i) const { return reinterpret_cast<const uint32_t *>(so->y)[i]; }
    BB_DEV void set_oxy(int k, int x, int y)
    {
        so->x[k] = (uint8_t)x; so->y[k] = (uint8_t)y; ot->x[k] = (uint8_t)x; ot->y[k] = (uint8_t)y;
    }
    BB_DEV uint32_t desc_mask(int d) const { return si->desc_mask[d]; }
    BB_DEV int leaf_kind(int l) const { return si->leaf_kind[l]; }
    BB_DEV int leaf_pre(int l) const { return si->leaf_pre[l]; }
    BB_DEV void set_leaf_pre(int l, int v) { si->leaf_pre[l] = (uint8_t)v; ins->leaf_pre[l] = (uint8_t)v; }
    BB_DEV int root_kind() const { return si->root_kind; }
    BB_DEV int side_and() const { return si->side_and; }
    BB_DEV void set_side_and(int v) { si->side_and = (uint8_t)v; ins->side_and = (uint8_t)v; }
    BB_DEV int flags() const { return si->flags; }
    BB_DEV void set_flags(int v) { si->flags = (uint8_t)v; ins->flags = (uint8_t)v; }
    BB_DEV int start_carry() const { return si->pad0; }
};

// ring slot -> live state (global) and -> the staged copy (shared), by the 8 lanes of the group together
template <class PP>
BB_DEV void swap_in8(const LevelParams &lp, const PP &P, int env, int slot, int r, uint8_t *srec)
{
    const LevelOut o = r2_ring_slot(lp, P, env, slot);
    const uint4 *sg = reinterpret_cast<const uint4 *>(o.grid);
    uint4 *lg = reinterpret_cast<uint4 *>(P.grid + (size_t)env * lp.cells_pad);
    uint4 *mg = reinterpret_cast<uint4 *>(srec);
    const int gch = lp.cells_pad / 16;
    for (int i = r; i < gch; i += 8) { const uint4 v = BB_LDCG(sg + i); lg[i] = v; mg[i] = v; }
    if (r < 6) { const uint4 v = BB_LDCG(reinterpret_cast<const uint4 *>(o.obj) + r); reinterpret_cast<uint4 *>(P.obj + env)[r] = v; mg[gch + r] = v; }
    if (r >= 5) { const uint4 v = BB_LDCG(reinterpret_cast<const uint4 *>(o.ins) + (r - 5)); reinterpret_cast<uint4 *>(P.ins + env)[r - 5] = v; mg[gch + 6 + (r - 5)] = v; }
    const uint4 *st = reinterpret_cast<const uint4 *>(o.tok);
    uint4 *lt = reinterpret_cast<uint4 *>(P.tok + (size_t)env * lp.max_tokens);
    for (int i = r; i < lp.max_tokens / 8; i += 8) lt[i] = BB_LDCG(st + i);
}

// UNTR: KIND_UNLOCK pools (objects without a table entry, env_logic.cuh CARRY_UNTRACKED); every other level runs the
// UNTR = false instantiations
template <class PP, int ACT_BYTES, bool UNTR>
BB_DEV void step8_role(const LevelParams &lp, const PP &P, const void *__restrict__ actions, uint8_t *__restrict__ obs,
                       float *__restrict__ reward, uint8_t *__restrict__ done, int8_t *__restrict__ dirs,
                       const int n, const int mode, const int force_reset, uint8_t *smem8, const int lane, const int warp, const unsigned cta)
{
    // smem8: [16 envs][cells_pad + 144] then the tiles
    const int r = lane & 7, g = lane >> 3;
    const int wg = cta * S8_WARPS + warp;
    const int env0 = wg * 4, env = env0 + g;
    const bool valid = env < n;
    const int leader = lane & ~7;
    const unsigned gmask = 0xFFu << (8 * g);
    const int rec_bytes = lp.cells_pad + S8_REC_FIXED;
    uint8_t *srec = smem8 + (size_t)(warp * 4 + g) * rec_bytes;    // this env's staged record
    uint32_t *tile = reinterpret_cast<uint32_t *>(smem8 + (size_t)S8_WARPS * 4 * rec_bytes) + warp * (S8_TILE_WORDS + 1);
    const int gch = lp.cells_pad / 16;

    // ---- all of the env's state in flight at once: asynchronous copies to shared memory + the hot record
    const size_t e = (size_t)(valid ? env : 0);
    if (valid) {
        const uint4 *gsrc = reinterpret_cast<const uint4 *>(P.grid + e * lp.cells_pad);
        for (int i = r; i < gch; i += 8) BB_CP_ASYNC16(srec + 16 * i, gsrc + i);
        if (r < 6) BB_CP_ASYNC16(srec + 16 * (gch + r), reinterpret_cast<const uint4 *>(P.obj + e) + r);
        if (r >= 5) BB_CP_ASYNC16(srec + 16 * (gch + 6 + r - 5), reinterpret_cast<const uint4 *>(P.ins + e) + (r - 5));
    }
    EnvHot h;
    { uint4 z = make_uint4(0, 0, 0, 0); h = *reinterpret_cast<EnvHot *>(&z); }
    int a = 0;
    if (valid) {
        h = P.hot[env];
        if (r == 0 && !force_reset) {
            if (ACT_BYTES == 1) a = reinterpret_cast<const int8_t *>(actions)[env];
            else a = (int)reinterpret_cast<const long long *>(actions)[env];
        }
    }
    BB_CP_ASYNC_WAIT_ALL();
    BB_SYNCWARP();
    StagedMem mem(lp, srec, reinterpret_cast<ObjTab *>(srec + lp.cells_pad), reinterpret_cast<InstrRec *>(srec + lp.cells_pad + sizeof(ObjTab)),
                  P.grid + e * lp.cells_pad, P.obj + e, P.ins + e);
    bool stepped = false, ended = false, succeeded = false, error = false, begin = force_reset != 0;
    float rew = 0.0f; bool dn = false;
    if (valid && r == 0 && !force_reset) {                        // the group's leader steps the env
        if (!(h.dirflags & 4)) {
            const StepResult sr = step_env<UNTR>(h, mem, a);
            rew = sr.reward; dn = sr.done;
            stepped = true; ended = dn; succeeded = sr.success;
            if (dn) {
                if (mode == BB_MODE_AUTORESET) begin = true;
                else { h.dirflags |= 4; P.last_reward[env] = rew; }
            }
        } else { rew = P.last_reward[env]; dn = true; }           // ManyEnvs: replay the last result
    }
    BB_SYNCWARP();
    {   // leader's pose and "episode begins" flag to the whole group
        uint4 hv = *reinterpret_cast<uint4 *>(&h);
        hv.x = BB_SHFL(hv.x, leader); hv.y = BB_SHFL(hv.y, leader);
        hv.z = BB_SHFL(hv.z, leader); hv.w = BB_SHFL(hv.w, leader);
        h = *reinterpret_cast<EnvHot *>(&hv);
        begin = BB_SHFL(begin ? 1 : 0, leader) != 0;
    }
    if (begin && valid) {                                         // uniform within the group
        const uint32_t hd = P.head[env];
        const uint32_t tl = BB_LDCG(P.tail_pub + env);
        if (tl - hd >= 1u && tl - hd <= (uint32_t)P.depth) {
            const int slot = (int)(hd % (uint32_t)P.depth);
            swap_in8(lp, P, env, slot, r, srec);                  // the 8 lanes copy the level together
            const uint4 hv = BB_LDCG(reinterpret_cast<const uint4 *>(r2_ring_slot(lp, P, env, slot).hot));
            h = *reinterpret_cast<const EnvHot *>(&hv);
            BB_SYNCWARP_MASK(gmask);
            if (r == 0) P.head[env] = hd + 1u;
        } else if (r == 0) { error = true; *P.err_flag = 1; }     // ring dry (the host orders k_gen first): the next call fails
    }
    if (valid && r == 0) {
        P.hot[env] = h;
        if (reward) reward[env] = rew;
        if (done) done[env] = dn ? 1 : 0;
        if (dirs) dirs[env] = (int8_t)(h.dirflags & 3);
    }
    BB_SYNCWARP();                                                 // staged-copy writes above are visible to the column loads
    // ---- observation: lane r < 7 holds view column vi = r --------------------------------------
    const ViewGeom v = view_geom(lp, h.x, h.y, h.dirflags & 3);
    uint32_t lo = 0, hi = 0, cm = 0;
    if (valid && r < 7) { col_load(mem, v, r, lo, hi); cm = col_see(lo, hi); }
    uint32_t see[7], vis[7];
#pragma unroll
    for (int j = 0; j < 7; j++) see[j] = (BB_BALLOT((cm >> j) & 1u) >> (8 * g)) & 0x7Fu;
    vis_rows(see, vis);
    uint32_t cv = 0;
#pragma unroll
    for (int j = 0; j < 7; j++) cv |= ((vis[j] >> r) & 1u) << j;
    if (r == 3 && valid) hi = (hi & 0xFF00FFFFu) | ((uint32_t)carry_cell_of<UNTR>(h, mem) << 16);   // own cell: what it carries
    uint32_t o[6];
    col_encode(lo, hi, (valid && r < 7) ? cv : 0u, o);
    // ---- stage 28 records of 21 bytes, then coalesced stores ------------------------------------
    const uint32_t next_w0 = BB_SHFL(o[0], r < 6 ? lane + 1 : lane + 2);
    if (r < 7) stage_record_words<21, 6>(tile, o, 7 * g + r, next_w0);
    BB_SYNCWARP();
    int nv = n - env0; nv = nv > 4 ? 4 : nv;
    if (nv > 0) {
        uint8_t *dst = obs + (size_t)env0 * OBS_BYTES;
        if (nv == 4 && (((uintptr_t)dst) & 3) == 0) {
            uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
#pragma unroll
            for (int i = 0; i < (S8_TILE_WORDS + 31) / 32; i++) { const int idx = lane + 32 * i; if (idx < S8_TILE_WORDS) d32[idx] = tile[idx]; }
        } else {
            const uint8_t *sb = reinterpret_cast<const uint8_t *>(tile);
            for (int i = lane; i < nv * OBS_BYTES; i += 32) dst[i] = sb[i];
        }
    }
    // ---- counters: one slot per warp, reductions without a return value (no stall at exit) ------
    const uint32_t m_step = BB_BALLOT(stepped), m_end = BB_BALLOT(ended);
    const uint32_t m_succ = BB_BALLOT(succeeded), m_err = BB_BALLOT(error);
    if (lane == 0) {
        unsigned long long *c = P.warp_counters + 4ull * wg;
        if (m_step) BB_ATOMIC_ADD(c + 0, (unsigned long long)BB_POPC(m_step));
        if (m_end) BB_ATOMIC_ADD(c + 1, (unsigned long long)BB_POPC(m_end));
        if (m_succ) BB_ATOMIC_ADD(c + 2, (unsigned long long)BB_POPC(m_succ));
        if (m_err) BB_ATOMIC_ADD(c + 3, (unsigned long long)BB_POPC(m_err));
    }
}

}  // namespace bb
