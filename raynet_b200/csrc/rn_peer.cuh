// rn_peer.cuh -- the one exchange step of the path as ONE kernel over NVLink peer memory.
//
// After a sweep every rank holds a PARTIAL occupancy accumulator (the sum of its own rays' messages).  NCCL's
// all-reduce of the 64 MiB grid (C3) plus the fill of the next partial cost ~0.4 ms per sweep; on the fixed-size job
// sharded over 8 GPUs a whole sweep is ~0.5 ms, so the exchange decides the strong-scaling efficiency.  This kernel
// does it over peer-mapped buffers (torch symmetric memory provides the mapping; NVSwitch gives every GPU full
// bandwidth to every peer):
//   barrier   every CTA b tells CTA b of every peer that its rank's sweep is complete (release store into the
//             peer's flag array, acquire spin on its own);
//   reduce + broadcast   rank r owns slice r of the grid: it loads that slice from all N partials (P2P loads,
//             128-bit, 2 N in flight per thread), adds them and the prior, and stores the result into slice r of all
//             N result buffers (P2P stores) -- each element crosses every link exactly once in each direction;
//   zero      the partial of the NEXT sweep (double-buffered: nobody has touched it since the previous exchange
//             completed on every rank), so no separate fill launch sits between the sweeps;
//   barrier   results complete everywhere, and every peer is done reading this rank's partial.
// Fused into the exchange: the prior (no seeded rank, no epilogue pass) and the fill.  The grid is one CTA per SM so
// that all CTAs of all ranks are resident (a CTA only ever waits for its twin on the other GPUs).
#pragma once

#include "rn_common.cuh"

#define RN_PEER_MAX_WORLD 16

struct PeerArgs {
    const float *partial[RN_PEER_MAX_WORLD];   // every rank's partial accumulator of THIS sweep (peer-mapped), [n]
    float *result[RN_PEER_MAX_WORLD];          // every rank's result buffer (peer-mapped), [n]
    uint32_t *flags[RN_PEER_MAX_WORLD];        // every rank's flag array (peer-mapped), [gridDim.x][world]
    float *zero;                               // local: the partial of the next sweep, cleared here (may be null)
    int rank, world;
    uint32_t epoch;                            // this call uses epoch + 1 and epoch + 2
    float prior;
    int64_t n;                                 // elements (multiple of 4)
    const float *mc_partial;                   // multicast (NVLS) address of the partials / results: one load is reduced
    float *mc_result;                          //   by the switch over all ranks, one store lands on all ranks
};

__device__ __forceinline__ void rn_st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t rn_ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 rn_ld_peer4(const float4 *p) {   // never from a (possibly stale) L1 line
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// CTA b of this rank <-> CTA b of every peer.  Everything this CTA wrote before is visible to the peers' CTA b
// after they pass; flags only ever grow (wrap-safe comparison).
__device__ __forceinline__ void rn_peer_barrier(const PeerArgs &a, uint32_t value) {
    __syncthreads();
    if ((int)threadIdx.x < a.world) {
        const int peer = threadIdx.x;
        __threadfence_system();
        rn_st_release_sys(a.flags[peer] + (size_t)blockIdx.x * a.world + a.rank, value);
        const uint32_t *mine = a.flags[a.rank] + (size_t)blockIdx.x * a.world + peer;
        while ((int32_t)(rn_ld_acquire_sys(mine) - value) < 0) { }
    }
    __syncthreads();
}

template <int WORLD>   // 0: run-time world size
__global__ void __launch_bounds__(512) peer_allreduce_kernel(PeerArgs a) {
    rn_peer_barrier(a, a.epoch + 1);
    const int world = WORLD ? WORLD : a.world;
    const int64_t n4 = a.n >> 2;
    const int64_t per = (n4 + world - 1) / world;
    const int64_t lo = per * a.rank, hi = min(n4, lo + per);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (WORLD) {
        const float4 *src[WORLD ? WORLD : 1];
        float4 *dst[WORLD ? WORLD : 1];
#pragma unroll
        for (int p = 0; p < WORLD; p++) {   // peers staggered by rank so that the links are loaded evenly
            const int q = (a.rank + p) % (WORLD ? WORLD : 1);
            src[p] = reinterpret_cast<const float4 *>(a.partial[q]);
            dst[p] = reinterpret_cast<float4 *>(a.result[q]);
        }
        for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += 2 * stride) {
            const int64_t i2 = i + stride;
            const bool two = i2 < hi;
            float4 v[WORLD ? WORLD : 1], w[WORLD ? WORLD : 1];
#pragma unroll
            for (int p = 0; p < WORLD; p++) {
                v[p] = rn_ld_peer4(src[p] + i);
                w[p] = two ? rn_ld_peer4(src[p] + i2) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float4 s = make_float4(a.prior, a.prior, a.prior, a.prior), t = s;
#pragma unroll
            for (int p = 0; p < WORLD; p++) {
                s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w;
                t.x += w[p].x; t.y += w[p].y; t.z += w[p].z; t.w += w[p].w;
            }
#pragma unroll
            for (int p = 0; p < WORLD; p++) {
                dst[p][i] = s;
                if (two) dst[p][i2] = t;
            }
        }
    } else {
        for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
            float4 s = make_float4(a.prior, a.prior, a.prior, a.prior);
            for (int p = 0; p < world; p++) {
                const float4 v = rn_ld_peer4(reinterpret_cast<const float4 *>(a.partial[(a.rank + p) % world]) + i);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            for (int p = 0; p < world; p++) reinterpret_cast<float4 *>(a.result[(a.rank + p) % world])[i] = s;
        }
    }
    if (a.zero) {
        float4 *z = reinterpret_cast<float4 *>(a.zero);
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) z[i] = zero4;
    }
    rn_peer_barrier(a, a.epoch + 2);
}

// ---- the same exchange with the reduction and the broadcast done INSIDE the NVSwitch (NVLS) ---------------------------
// multimem.ld_reduce on the multicast address of the partials returns the sum over all ranks' copies, multimem.st on
// the multicast address of the results writes all ranks' copies: a rank pulls its slice once (n / world elements in)
// and pushes it once (n / world out) instead of world - 1 times each; every link carries n elements per direction
// instead of 2 n (world - 1) / world.  Barriers, prior and fill as above.  The order in which the switch adds the
// world operands is not specified: float32 sums differ from the peer-load kernel's in the last bit.
__device__ __forceinline__ float4 rn_mc_ld_reduce4(const float *mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void rn_mc_st4(float *mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

#ifndef RN_PEER_MC_UNROLL
#define RN_PEER_MC_UNROLL 4
#endif
__global__ void __launch_bounds__(512) peer_allreduce_mc_kernel(PeerArgs a) {
    rn_peer_barrier(a, a.epoch + 1);
    const int64_t n4 = a.n >> 2;
    const int64_t per = (n4 + a.world - 1) / a.world;
    const int64_t lo = per * a.rank, hi = min(n4, lo + per);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += RN_PEER_MC_UNROLL * stride) {
        float4 v[RN_PEER_MC_UNROLL];
#pragma unroll
        for (int u = 0; u < RN_PEER_MC_UNROLL; u++)
            if (i + u * stride < hi) v[u] = rn_mc_ld_reduce4(a.mc_partial + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < RN_PEER_MC_UNROLL; u++)
            if (i + u * stride < hi) {
                v[u].x += a.prior; v[u].y += a.prior; v[u].z += a.prior; v[u].w += a.prior;
                rn_mc_st4(a.mc_result + 4 * (i + u * stride), v[u]);
            }
    }
    if (a.zero) {
        float4 *z = reinterpret_cast<float4 *>(a.zero);
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) z[i] = zero4;
    }
    rn_peer_barrier(a, a.epoch + 2);
}
