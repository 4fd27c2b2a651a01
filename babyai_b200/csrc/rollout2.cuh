// rollout2.cuh -- the stepping role of k_rollout2 (pool.cu): the persistent rollout kernel with TWO LANES PER ENVIRONMENT.
// EXPERIMENTAL (BB_ROLLOUT_LANES=2): written after round 1's GPU budget was spent, never run on a GPU yet.
//
// k_rollout is latency bound with 4.2 warps per scheduler and 8.7 cycles per issued instruction
// (profiles/r01y_ncu_rollout_details.txt); the number of stepping warps is capped by the problem size (65 536 envs =
// 2 048 warps), so the remaining parallelism is inside an env.  Here a warp serves 16 envs:
//   * the even lane of a pair applies the action and runs the verifier (step_env) and broadcasts the hot record;
//   * a finished env's next level is copied from the ring by the two lanes together (half the dependent loads each);
//   * the observation is split by view columns (even lane: columns 0..3 = 84 bytes, odd lane: 4..6 = 63 bytes; one
//     shuffle exchanges the see-through bits on multi-room levels), encoded per column and staged as 21-byte records
//     (pair_cols_load / pair_cols_encode / pair_stage in env_logic.cuh).
// Twice the stepping warps for an estimated 1.3-1.5 x the instructions.  CTA = 4 stepping warps (64 envs, the same
// shared-memory footprint per env as k_rollout) + the generator warp of the fused mode; 7 CTAs per SM = 35 warps need
// <= 58 registers per thread (compiles to 56 with ~90 bytes of spills).
//
// The warp primitives are macros so that tests/hostemu can compile this very function for the host with one OS thread
// per lane (tests/hostemu/simt_rollout2.cpp): the GPU-less suite runs whole rollouts through it against the per-step path.
#pragma once
#include "../../include/babyai_b200.h"
#include "env_logic.cuh"

#if defined(__CUDACC__)
#define BB_DEV __device__ __forceinline__
#define BB_SYNCWARP() __syncwarp()
#define BB_SYNCTHREADS() __syncthreads()
#define BB_SHFL(v, src) __shfl_sync(0xFFFFFFFFu, (v), (src))
#define BB_SHFL_XOR(v, m) __shfl_xor_sync(0xFFFFFFFFu, (v), (m))
#define BB_SHFL_DOWN(v, d) __shfl_down_sync(0xFFFFFFFFu, (v), (d))
#define BB_LDCG(p) __ldcg(p)
#define BB_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define BB_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
// sign-extending byte load (no dependent conversion instruction after the load)
static __device__ __forceinline__ int bb_ld_s8(const int8_t *p) { int v; asm volatile("ld.global.nc.s8 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
#define BB_LD_S8(p) bb_ld_s8(p)
#endif

namespace bb {

constexpr int R2_WARPS = 4, R2_ENVS = 16;                        // stepping warps per CTA, envs per warp
constexpr int R2_THREADS = 32 * R2_WARPS, R2_THREADS_FUSED = R2_THREADS + 32;
constexpr int TILE2_WORDS = R2_ENVS * OBS_BYTES / 4;              // 588 words = 2352 B per warp
constexpr int R2_OBJ_STRIDE = 25, R2_INS_STRIDE = 13;            // odd word strides of the shared-memory records (= pool.cu SM_*_STRIDE)

template <class PP>
BB_DEV LevelOut r2_ring_slot(const LevelParams &lp, const PP &P, int env, int slot)
{
    const size_t idx = (size_t)slot * P.n + env;
    LevelOut o;
    o.grid = P.rgrid + idx * lp.cells_pad; o.hot = P.rhot + idx; o.obj = P.robj + idx; o.ins = P.rins + idx;
    o.tok = P.rtok + idx * lp.max_tokens;
    return o;
}

// coalesced copy of `chunks_per_env` 16-byte chunks per env between global memory (contiguous records of the warp's envs)
// and the env-strided shared-memory records
template <bool TO_SMEM>
BB_DEV void r2_copy_records(uint32_t *sm, int stride_words, uint4 *glob, int chunks_per_env, int nv, int lane)
{
    for (int idx = lane; idx < nv * chunks_per_env; idx += 32) {
        const int e = idx / chunks_per_env, w0 = (idx - e * chunks_per_env) * 4;
        uint32_t *d = sm + e * stride_words + w0;
        if (TO_SMEM) { const uint4 v = glob[idx]; d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
        else glob[idx] = make_uint4(d[0], d[1], d[2], d[3]);
    }
}

BB_DEV void store_tile2(const uint32_t *tile, uint8_t *dst, int lane, int valid_envs)
{
    if (valid_envs == R2_ENVS && (((uintptr_t)dst) & 15) == 0) {
        const uint4 *s = reinterpret_cast<const uint4 *>(tile);
        uint4 *d = reinterpret_cast<uint4 *>(dst);
#pragma unroll
        for (int i = 0; i < (TILE2_WORDS / 4 + 31) / 32; i++) {
            const int idx = lane + 32 * i;
            if (idx < TILE2_WORDS / 4) d[idx] = s[idx];
        }
    } else {                                                  // ragged tail / unaligned destination
        const uint8_t *s = reinterpret_cast<const uint8_t *>(tile);
        const int nbytes = valid_envs * OBS_BYTES;
        for (int i = lane; i < nbytes; i += 32) dst[i] = s[i];
    }
}

// One stepping warp: 16 envs, lane = 2 * (env within the warp) + half.  warp_smem: the warp's shared-memory area
// (R2_ENVS records of grid / object table / verifier record + the observation tile); warp_global: its index in the grid.
// UNTR: KIND_UNLOCK pools (untracked objects, env_logic.cuh CARRY_UNTRACKED)
template <class PP, class MemT, bool UNTR = false>
BB_DEV void rollout2_step_warp(const LevelParams &lp, const PP &P, const int8_t *actions, uint8_t *obs, float *reward, uint8_t *done,
                               int8_t *dirs, const int n, const int T, const int mode, const bool fused, uint32_t *warp_smem,
                               const int lane, const int warp_global, volatile int *s_done)
{
    const int gs = (lp.cells_pad >> 2) | 1;
    const int el = lane >> 1, hf = lane & 1, even = lane & ~1;
    const int env0 = warp_global * R2_ENVS, env = env0 + el;
    int nv = n - env0; nv = nv > R2_ENVS ? R2_ENVS : (nv < 0 ? 0 : nv);
    const bool valid = el < nv;
    uint32_t *sg = warp_smem, *so = sg + R2_ENVS * gs, *si = so + R2_ENVS * R2_OBJ_STRIDE;
    uint32_t *tile = si + R2_ENVS * R2_INS_STRIDE;          // 16-byte aligned: every term is a multiple of 4 words
    r2_copy_records<true>(sg, gs, reinterpret_cast<uint4 *>(P.grid + (size_t)env0 * lp.cells_pad), lp.cells_pad >> 4, nv, lane);
    r2_copy_records<true>(so, R2_OBJ_STRIDE, reinterpret_cast<uint4 *>(P.obj + env0), 6, nv, lane);
    r2_copy_records<true>(si, R2_INS_STRIDE, reinterpret_cast<uint4 *>(P.ins + env0), 3, nv, lane);
    EnvHot h;
    { uint4 z = make_uint4(0, 0, 0, 0); h = *reinterpret_cast<EnvHot *>(&z); }
    uint32_t head = 0, avail = 0;
    float last_rew = 0.0f;
    if (valid) {
        h = P.hot[env];
        head = P.head[env];
        avail = (fused ? P.tail[env] : BB_LDCG(P.tail_pub + env)) - head;
        if (mode == BB_MODE_FREEZE) last_rew = P.last_reward[env];
    }
    if (fused) BB_SYNCTHREADS();
    BB_SYNCWARP();
    MemT mem(lp, reinterpret_cast<uint8_t *>(sg + el * gs), reinterpret_cast<uint8_t *>(so + el * R2_OBJ_STRIDE),
                    reinterpret_cast<uint8_t *>(si + el * R2_INS_STRIDE));
    const bool single_room = lp.num_rows == 1 && lp.num_cols == 1;
    uint32_t n_step = 0, n_end = 0, n_succ = 0, n_err = 0, consumed = 0;      // counters: even lanes only
    int a_next = 0;
    if (valid) a_next = BB_LD_S8(actions + env);
    for (int t = 0; t < T; t++) {
        const int a = a_next;
        if (valid && t + 1 < T) a_next = BB_LD_S8(actions + (size_t)(t + 1) * n + env);
        float rew = 0.0f; bool dn = false; int begin = 0;
        if (valid && hf == 0) {                             // the even lane steps the env
            if (!(h.dirflags & 4)) {
                const StepResult sr = step_env<UNTR>(h, mem, a);
                rew = sr.reward; dn = sr.done;
                n_step++; n_end += dn; n_succ += sr.success;
                if (dn) {
                    if (mode == BB_MODE_AUTORESET) begin = 1;
                    else { h.dirflags |= 4; last_rew = rew; }
                }
            } else { rew = last_rew; dn = true; }
        }
        BB_SYNCWARP();                                       // step_env's shared-memory writes -> the partner lane
        {   // the pair's hot record and "episode begins" flag from the even lane
            uint4 hv = *reinterpret_cast<uint4 *>(&h);
            hv.x = BB_SHFL(hv.x, even); hv.y = BB_SHFL(hv.y, even);
            hv.z = BB_SHFL(hv.z, even); hv.w = BB_SHFL(hv.w, even);
            h = *reinterpret_cast<EnvHot *>(&hv);
            begin = BB_SHFL(begin, even);
        }
        if (begin && valid) {                               // uniform within the pair: both lanes copy the next level
            if (consumed < avail && avail <= (uint32_t)P.depth) {
                const LevelOut o = r2_ring_slot(lp, P, env, (int)((head + consumed) % (uint32_t)P.depth));
                uint32_t *mg = reinterpret_cast<uint32_t *>(mem.g);
                for (int k = hf; k < lp.cells_pad / 16; k += 2) {
                    const uint4 v = BB_LDCG(reinterpret_cast<const uint4 *>(o.grid) + k);
                    mg[4 * k] = v.x; mg[4 * k + 1] = v.y; mg[4 * k + 2] = v.z; mg[4 * k + 3] = v.w;
                }
                uint32_t *mo = reinterpret_cast<uint32_t *>(mem.o);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const int c = 2 * k + hf;
                    const uint4 v = BB_LDCG(reinterpret_cast<const uint4 *>(o.obj) + c);
                    mo[4 * c] = v.x; mo[4 * c + 1] = v.y; mo[4 * c + 2] = v.z; mo[4 * c + 3] = v.w;
                }
                uint32_t *mi = reinterpret_cast<uint32_t *>(mem.i);
                for (int c = hf; c < 3; c += 2) {
                    const uint4 v = BB_LDCG(reinterpret_cast<const uint4 *>(o.ins) + c);
                    mi[4 * c] = v.x; mi[4 * c + 1] = v.y; mi[4 * c + 2] = v.z; mi[4 * c + 3] = v.w;
                }
                uint4 *lt = reinterpret_cast<uint4 *>(P.tok + (size_t)env * lp.max_tokens);
                for (int k = hf; k < lp.max_tokens / 8; k += 2) lt[k] = BB_LDCG(reinterpret_cast<const uint4 *>(o.tok) + k);
                const uint4 hv = BB_LDCG(reinterpret_cast<const uint4 *>(o.hot));
                h = *reinterpret_cast<const EnvHot *>(&hv);
                consumed++;
            } else if (hf == 0) n_err++;
        }
        BB_SYNCWARP();                                       // the swapped-in level -> both lanes
        if (valid && hf == 0 && mode == BB_MODE_AUTORESET && (int)h.step_count + 2 == (int)h.max_steps && consumed < avail) {
            const LevelOut o = r2_ring_slot(lp, P, env, (int)((head + consumed) % (uint32_t)P.depth));
            BB_PREFETCH_L2(o.grid);
            BB_PREFETCH_L2(o.obj);
            BB_PREFETCH_L2(reinterpret_cast<const uint8_t *>(o.obj) + 64);
            BB_PREFETCH_L2(o.ins);
            BB_PREFETCH_L2(reinterpret_cast<const uint8_t *>(o.ins) + 32);
            BB_PREFETCH_L2(o.hot);
            BB_PREFETCH_L2(o.tok);
        }
        // ---- observation: this lane's half of the view columns ----
        uint32_t lo[4], hi[4], oc[4][6];
        const int dir = h.dirflags & 3;
        const ViewGeom v = view_geom(lp, h.x, h.y, dir);
        uint32_t cm = 0;
        if (valid) cm = pair_cols_load(mem, v, hf, lo, hi);
        uint32_t other = 0;
        if (!single_room) other = BB_SHFL_XOR(cm, 1);
        if (valid) pair_cols_encode(lp, v, h.x, h.y, dir, carry_cell_of<UNTR>(h, mem), hf, lo, hi, hf ? other : cm, hf ? cm : other, oc);
        else {
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int j = 0; j < 6; j++) oc[k][j] = 0;
        }
        const uint32_t next_first = BB_SHFL_DOWN(oc[0][0], 1);
        pair_stage(tile, oc, el, hf, next_first);
        if (valid && hf == 0) {
            const size_t oi = (size_t)t * n + env;
            if (reward) reward[oi] = rew;
            if (done) done[oi] = dn ? 1 : 0;
            if (dirs) dirs[oi] = (int8_t)dir;
        }
        BB_SYNCWARP();
        if (nv > 0) store_tile2(tile, obs + ((size_t)t * n + env0) * OBS_BYTES, lane, nv);
        BB_SYNCWARP();                                       // the tile is rewritten in the next iteration
    }
    BB_SYNCWARP();
    r2_copy_records<false>(sg, gs, reinterpret_cast<uint4 *>(P.grid + (size_t)env0 * lp.cells_pad), lp.cells_pad >> 4, nv, lane);
    r2_copy_records<false>(so, R2_OBJ_STRIDE, reinterpret_cast<uint4 *>(P.obj + env0), 6, nv, lane);
    r2_copy_records<false>(si, R2_INS_STRIDE, reinterpret_cast<uint4 *>(P.ins + env0), 3, nv, lane);
    if (valid && hf == 0) {
        P.hot[env] = h;
        P.head[env] = head + consumed;
        if (mode == BB_MODE_FREEZE) P.last_reward[env] = last_rew;
    }
    for (int off = 16; off; off >>= 1) {
        n_step += BB_SHFL_DOWN(n_step, off); n_end += BB_SHFL_DOWN(n_end, off);
        n_succ += BB_SHFL_DOWN(n_succ, off); n_err += BB_SHFL_DOWN(n_err, off);
    }
    if (fused && lane == 0) BB_ATOMIC_ADD(const_cast<int *>(s_done), 1);
    if (lane == 0) {
        unsigned long long *c = P.warp_counters + 4ull * warp_global;
        if (n_step) BB_ATOMIC_ADD(c + 0, (unsigned long long)n_step);
        if (n_end) BB_ATOMIC_ADD(c + 1, (unsigned long long)n_end);
        if (n_succ) BB_ATOMIC_ADD(c + 2, (unsigned long long)n_succ);
        if (n_err) BB_ATOMIC_ADD(c + 3, (unsigned long long)n_err);
    }
}

}  // namespace bb
