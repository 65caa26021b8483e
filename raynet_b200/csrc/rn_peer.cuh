// rn_peer.cuh -- the one exchange step of the path as ONE kernel over NVLink peer memory.
//
// After a sweep every rank holds a PARTIAL occupancy accumulator (the sum of its own rays' messages).  NCCL's
// all-reduce of the 64 MiB grid (C3) takes ~0.28 ms per sweep; on the fixed-size job sharded over 8 GPUs a whole
// sweep is ~0.5 ms, so the collective decides the strong-scaling efficiency.  This kernel does the exchange
// itself over peer-mapped buffers (torch symmetric memory provides the mapping; NVSwitch gives every GPU
// full bandwidth to every peer):
//   barrier   every CTA b tells CTA b of every peer that its rank's sweep is complete (release store into the
//             peer's flag array, acquire spin on its own);
//   reduce + broadcast   rank r owns slice r of the grid: it loads that slice from all N partials (P2P loads,
//             128-bit), adds them and the prior, and stores the result into slice r of all N result buffers
//             (P2P stores) -- each element crosses every link exactly once in each direction;
//   barrier   results complete everywhere, and every peer is done reading this rank's partial.
// Fused into the exchange: the prior (no seeded rank, no epilogue pass).  The grid is one CTA per SM so that all
// CTAs of all ranks are resident (a CTA only ever waits for its twin on the other GPUs).
#pragma once

#include "rn_common.cuh"

#define RN_PEER_MAX_WORLD 16

struct PeerArgs {
    const float *partial[RN_PEER_MAX_WORLD];   // every rank's partial accumulator (peer-mapped), [n]
    float *result[RN_PEER_MAX_WORLD];          // every rank's result buffer (peer-mapped), [n]
    uint32_t *flags[RN_PEER_MAX_WORLD];        // every rank's flag array (peer-mapped), [gridDim.x][world]
    int rank, world;
    uint32_t epoch;                            // this call uses epoch + 1 and epoch + 2
    float prior;
    int64_t n;                                 // elements (multiple of 4)
};

__device__ __forceinline__ void rn_st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t rn_ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
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

__global__ void __launch_bounds__(512) peer_allreduce_kernel(PeerArgs a) {
    rn_peer_barrier(a, a.epoch + 1);
    const int64_t n4 = a.n >> 2;
    const int64_t per = (n4 + a.world - 1) / a.world;
    const int64_t lo = per * a.rank, hi = min(n4, lo + per);
    const float4 pr = make_float4(a.prior, a.prior, a.prior, a.prior);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
        float4 s = pr;
#pragma unroll 8
        for (int p = 0; p < a.world; p++) {
            // own partial first would not matter: every operand is read exactly once
            const int q = (a.rank + p) % a.world;     // stagger the peers so that the links are loaded evenly
            const float4 v = __ldcv(reinterpret_cast<const float4 *>(a.partial[q]) + i);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
#pragma unroll 8
        for (int p = 0; p < a.world; p++) {
            const int q = (a.rank + p) % a.world;
            reinterpret_cast<float4 *>(a.result[q])[i] = s;
        }
    }
    rn_peer_barrier(a, a.epoch + 2);
}
