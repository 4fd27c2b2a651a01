// simt.cuh -- the warp / CTA primitives of the rollout kernels as macros, so that tests/hostemu can compile the very
// functions the kernels call (rollout_cta.cuh, gen_round.cuh) for the host with ONE OS THREAD PER LANE
// (tests/hostemu/simt_rollout.cpp: shuffles, votes and barriers become rendezvous on a per-warp / per-CTA barrier).
// The device build maps them to the intrinsics.
#pragma once
#include "../../include/babyai_b200.h"
#include "env_logic.cuh"

#if defined(__CUDACC__)
#define BB_DEV __device__ __forceinline__
#define BB_SYNCWARP() __syncwarp()
#define BB_SYNCTHREADS() __syncthreads()
#define BB_SYNCWARP_MASK(m) __syncwarp(m)
// asynchronous 16-byte copy global -> shared (cp.async.cg: L2 only), and "all my copies have landed"
static __device__ __forceinline__ void bb_cp_async16(void *smem_dst, const void *gmem_src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
static __device__ __forceinline__ void bb_cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
#define BB_CP_ASYNC16(dst, src) bb_cp_async16((dst), (src))
#define BB_CP_ASYNC_WAIT_ALL() bb_cp_async_wait_all()
#define BB_SHFL(v, src) __shfl_sync(0xFFFFFFFFu, (v), (src))
#define BB_SHFL_XOR(v, m) __shfl_xor_sync(0xFFFFFFFFu, (v), (m))
#define BB_SHFL_DOWN(v, d) __shfl_down_sync(0xFFFFFFFFu, (v), (d))
#define BB_LDCG(p) __ldcg(p)
#define BB_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define BB_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
// sign-extending byte load (no dependent conversion instruction after the load)
static __device__ __forceinline__ int bb_ld_s8(const int8_t *p) { int v; asm volatile("ld.global.nc.s8 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
#define BB_LD_S8(p) bb_ld_s8(p)
#define BB_ANY(x) __any_sync(0xFFFFFFFFu, (x))
#define BB_BALLOT(x) __ballot_sync(0xFFFFFFFFu, (x))
#define BB_POPC(x) __popc(x)
#define BB_SYNCTHREADS_OR(x) __syncthreads_or(x)
// the rendezvous of a kernel's warp ROLES (stepping warps / generator warp call it from different places): a named
// barrier with an explicit thread count -- __syncthreads() is only defined when every thread reaches the same call
#define BB_ROLE_SYNC(nthreads) asm volatile("barrier.sync 1, %0;" ::"r"(nthreads) : "memory")
// shared -> global bulk copy on the async proxy (SASS UBLKCP): the writers of the shared-memory tile fence the proxy, one
// elected thread issues the copy and commits it as a bulk group; wait_group.read returns once the source may be rewritten
#define BB_FENCE_ASYNC_SMEM() asm volatile("fence.proxy.async.shared::cta;" ::: "memory")
static __device__ __forceinline__ void bb_bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\ncp.async.bulk.commit_group;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
#define BB_BULK_STORE(gdst, ssrc, bytes) bb_bulk_store((gdst), (ssrc), (bytes))
#define BB_BULK_WAIT_READ() asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory")
#define BB_BULK_WAIT_READ_N(n) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(n) : "memory")
// producer / consumer named barriers (ids are immediates: a register id makes ptxas reserve all 16 barriers): every thread
// of the `nthreads` that take part either arrives (does not wait) or syncs (waits for all of them)
#define BB_BAR_SYNC(id, nthreads) asm volatile("barrier.sync %0, %1;" ::"n"(id), "r"(nthreads) : "memory")
#define BB_BAR_ARRIVE(id, nthreads) asm volatile("barrier.arrive %0, %1;" ::"n"(id), "r"(nthreads) : "memory")
#endif

namespace bb {

template <class PP>
BB_DEV LevelOut r2_ring_slot(const LevelParams &lp, const PP &P, int env, int slot)
{
    const size_t idx = (size_t)slot * P.n + env;
    LevelOut o;
    o.grid = P.rgrid + idx * lp.cells_pad; o.hot = P.rhot + idx; o.obj = P.robj + idx; o.ins = P.rins + idx;
    o.tok = P.rtok + idx * lp.max_tokens;
    return o;
}

}  // namespace bb
